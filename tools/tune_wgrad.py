"""Balances the wgrad work partition: measures every CTA's end time (nerf_debug_set_wgrad_profile), attributes it to the
unit of the CTA's first segment, rescales the unit costs and repeats.  Prints the cost table to hard-code.
usage: python tools/tune_wgrad.py [rays=4096] [samples=192] [iters=6]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch_nerf_b200 as tn
lib = tn._lib.load()
P, VP = tn._lib.ptr, tn._lib.c_void_p
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
s = int(sys.argv[2]) if len(sys.argv) > 2 else 192
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 6
m = n * s
tiles = (m + 127) // 128
net = tn.NeRF(63, 27, precision="bf16").cuda()
packed = net.packed_weights(True)
ray_o = torch.randn(n, 3, device="cuda"); ray_d = torch.randn(n, 3, device="cuda")
t = torch.rand(n, s, device="cuda") * 4 + 2
sig = torch.empty(m, device="cuda"); rgb = torch.empty(m, 3, device="cuda")
cache = torch.empty(lib.nerf_mlp_bf16_cache_bytes(m), dtype=torch.uint8, device="cuda")
scratch = torch.empty(lib.nerf_mlp_bf16_bwd_scratch_bytes(m), dtype=torch.uint8, device="cuda")
tn._lib.check(lib.nerf_mlp_bf16_forward(P(packed, torch.uint8), None, None, P(ray_o), P(ray_d), P(t), s, m, P(sig), P(rgb),
                                        P(cache, torch.uint8), tn._lib.stream()), "fwd")
g_sigma = torch.randn(m, device="cuda") * 1e-3; g_rgb = torch.randn(m, 3, device="cuda") * 1e-3
grads = [torch.zeros_like(p) for p in net.parameters()]
gp = tn._lib.pointer_array(grads)
def bwd(mask):
    tn._lib.check(lib.nerf_mlp_bf16_backward_part(P(packed, torch.uint8), P(cache, torch.uint8), P(rgb), m, P(g_sigma), P(g_rgb), gp,
                                                  P(scratch, torch.uint8), mask, 0, tiles, 0, tn._lib.stream()), "bwd")
bwd(7); torch.cuda.synchronize()
def timed(reps=10):
    for _ in range(3): bwd(4)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): bwd(4)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
costs = np.array([64, 80, 80, 80, 80, 64, 80, 80, 80, 88, 84], dtype=np.float64) * 10  # built-in table x 10
prof = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
best = (1e9, None)
for it in range(iters):
    c_int = (ctypes.c_int * 11)(*[int(round(c)) for c in costs])
    lib.nerf_debug_set_wgrad_costs(c_int, 11)
    us = timed()
    lib.nerf_debug_set_wgrad_profile(VP(prof.data_ptr()))
    ends = []
    for _ in range(3):
        prof.zero_(); bwd(4); torch.cuda.synchronize()
        p = prof.cpu().view(148, 16).numpy()
        ends.append((p[:, 2] - p[:, 0].min()) / 1e3)
    lib.nerf_debug_set_wgrad_profile(None)
    end = np.median(np.stack(ends), 0)
    unit = p[:, 3]
    mean_end = end.mean()
    per_unit = np.array([end[unit == u].mean() if (unit == u).any() else mean_end for u in range(11)])
    print(f"iter {it}: wgrad {us:7.1f} us  CTA end min/mean/max {end.min():7.1f} {mean_end:7.1f} {end.max():7.1f}  costs {[int(round(c)) for c in costs]}")
    print("          per-unit mean end:", " ".join(f"{x:7.1f}" for x in per_unit))
    if us < best[0]: best = (us, [int(round(c)) for c in costs])
    costs = costs * (per_unit / mean_end) ** 1.0
print("best:", best)
lib.nerf_debug_set_wgrad_costs(None, 0)
