"""MMA issue-rate micro-benchmark + in-kernel timeline of the forward chain (prints a small report)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_nerf_b200 as tn

lib = tn._lib.load()
P, VP = tn._lib.ptr, tn._lib.c_void_p
out = torch.zeros(148, dtype=torch.int64, device="cuda")
for blocks in (148,):
    for mode, n in ((0, 256), (0, 128), (0, 64), (1, 256), (1, 128)):
        iters = 2000
        tn._lib.check(lib.nerf_selftest_mma_rate(blocks, iters, n, mode, VP(out.data_ptr()), tn._lib.stream()), "rate")
        torch.cuda.synchronize()
        cyc = out[:blocks].float()
        per = cyc / (iters * 4)
        print(f"mma_rate blocks={blocks:3d} mode={'SS' if mode == 0 else 'TS'} N={n:3d}: cycles/MMA mean {per.mean():.1f} max {per.max():.1f}  "
              f"(floor {128 * n / 256:.0f})")

n, s = 4096, 192
m = n * s
net = tn.NeRF(63, 27, precision="bf16").cuda()
packed = net.packed_weights()
ray_o = torch.randn(n, 3, device="cuda"); ray_d = torch.randn(n, 3, device="cuda")
t = torch.rand(n, s, device="cuda") * 4 + 2
sig = torch.empty(m, device="cuda"); rgb = torch.empty(m, 3, device="cuda")
tiles = 4
prof = torch.zeros(tiles * 10 * 8, dtype=torch.int64, device="cuda")
def fwd(cache=None):
    tn._lib.check(lib.nerf_mlp_bf16_forward(P(packed, torch.uint8), None, None, P(ray_o), P(ray_d), P(t), s, m, P(sig), P(rgb),
                                            P(cache, torch.uint8) if cache is not None else None, tn._lib.stream()), "fwd")
for label, cache in (("inference", None), ("training", torch.empty(lib.nerf_mlp_bf16_cache_bytes(m), dtype=torch.uint8, device="cuda"))):
    fwd(cache); torch.cuda.synchronize()
    lib.nerf_debug_set_profile_buffer(VP(prof.data_ptr()), tiles)
    prof.zero_()
    fwd(cache); torch.cuda.synchronize()
    lib.nerf_debug_set_profile_buffer(None, 0)
    p = prof.cpu().view(tiles, 10, 8)
    t0 = int(p[0, 0, 0])
    print(f"--- forward chain timeline ({label}), CTA 0, cycles relative to first layer start")
    print("tile layer  mma_start  mma_issued  acc_seen  epi_done | mma_span  epi_span  layer_period  wait_act  wait_w")
    prev = None
    for ti in range(1, tiles):
        for l in range(10):
            a, b, c, d = [int(x) - t0 for x in p[ti, l, :4]]
            period = (a - prev) if prev is not None else 0
            prev = a
            print(f"{ti:4d} {l:5d} {a:10d} {b:11d} {c:9d} {d:9d} | {b - a:8d} {d - c:9d} {period:9d} {int(p[ti, l, 4]):9d} {int(p[ti, l, 5]):7d}")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(5): fwd(cache)
    ev1.record(); torch.cuda.synchronize()
    print(f"{label}: {ev0.elapsed_time(ev1)/5*1e3:.0f} us per launch of {m} rows")
