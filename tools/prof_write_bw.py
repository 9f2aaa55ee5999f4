"""HBM write-bandwidth ceiling with different store flavours (see nerf_selftest_write_bw)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_nerf_b200 as tn
lib = tn._lib.load()
VP = tn._lib.c_void_p
nbytes = 4 << 30
x = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
names = {0: "st.global.v4", 1: "st.global.cs.v4", 2: "bulk store", 3: "bulk store evict_first", 4: "bulk store evict_last"}
for blocks in (148, 296, 592):
    for mode in range(5):
        def run(): tn._lib.check(tn._lib.load_selftest().nerf_selftest_write_bw(VP(x.data_ptr()), nbytes, mode, blocks, tn._lib.stream()), "wbw")
        run(); torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(3): run()
        ev1.record(); torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / 3
        print(f"blocks {blocks:4d} {names[mode]:24s}: {ms*1e3:7.0f} us -> {nbytes/ms/1e9:.2f} TB/s")
