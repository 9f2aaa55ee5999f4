"""Debug experiment: training-mode forward chain with the cache stores disabled / wrapped onto an L2-resident window."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_nerf_b200 as tn
lib = tn._lib.load()
P, VP = tn._lib.ptr, tn._lib.c_void_p
n, s = 4096, 192
m = n * s
net = tn.NeRF(63, 27, precision="bf16").cuda()
packed = net.packed_weights()
ray_o = torch.randn(n, 3, device="cuda"); ray_d = torch.randn(n, 3, device="cuda")
t = torch.rand(n, s, device="cuda") * 4 + 2
sig = torch.empty(m, device="cuda"); rgb = torch.empty(m, 3, device="cuda")
cache = torch.empty(lib.nerf_mlp_bf16_cache_bytes(m), dtype=torch.uint8, device="cuda")
def fwd(c):
    tn._lib.check(lib.nerf_mlp_bf16_forward(P(packed, torch.uint8), None, None, P(ray_o), P(ray_d), P(t), s, m, P(sig), P(rgb),
                                            P(c, torch.uint8) if c is not None else None, tn._lib.stream()), "fwd")
for mode, label in ((0, "normal"), (1, "no block stores"), (2, "stores wrapped onto 64 tiles"), (3, "no STS, no block stores"), (0, "normal again")):
    lib.nerf_debug_set_profile_buffer(None, mode << 16)
    fwd(cache); torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(5): fwd(cache)
    ev1.record(); torch.cuda.synchronize()
    print(f"training fwd [{label}]: {ev0.elapsed_time(ev1)/5*1e3:.0f} us")
lib.nerf_debug_set_profile_buffer(None, 0)
fwd(None); torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(5): fwd(None)
ev1.record(); torch.cuda.synchronize()
print(f"inference fwd: {ev0.elapsed_time(ev1)/5*1e3:.0f} us")
# pure write / read / copy bandwidth for reference
x = torch.empty(4 << 30, dtype=torch.uint8, device="cuda"); y = torch.empty(4 << 30, dtype=torch.uint8, device="cuda")
for label, fn, nbytes in (("fill (write only)", lambda: x.fill_(1), 4 << 30), ("copy (read+write)", lambda: y.copy_(x), 8 << 30),
                          ("sum (read only)", lambda: x.view(torch.int64).sum(), 4 << 30)):
    fn(); torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(5): fn()
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 5
    print(f"{label}: {ms*1e3:.0f} us -> {nbytes/ms/1e9:.2f} TB/s")
