"""Backward of one network as a two-stream pipeline: the weight-gradient kernel (HBM-bound) of tile range q runs on some
SMs while the activation-gradient chain (tensor / shared-memory bound) of range q+1 runs on the others.

Prints (1) both kernels alone at reduced grid sizes, (2) the pipelined schedule for several (parts, dgrad CTAs, wgrad
CTAs) against the sequential backward, with the gradients of every schedule checked against the sequential ones.
usage: python tools/prof_overlap.py [rays=4096] [samples=192]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_nerf_b200 as tn

lib = tn._lib.load()
P = tn._lib.ptr
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
s = int(sys.argv[2]) if len(sys.argv) > 2 else 192
m = n * s
tiles = (m + 127) // 128
torch.manual_seed(0)
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
grads = [torch.zeros_like(p) for p in net.ordered_parameters()]
gp = tn._lib.pointer_array(grads)
sA, sB = torch.cuda.Stream(), torch.cuda.Stream()


def part(phases, t0, t1, ctas, stream):
    tn._lib.check(lib.nerf_mlp_bf16_backward_part(P(packed, torch.uint8), P(cache, torch.uint8), P(rgb), m, P(g_sigma), P(g_rgb),
                                                  gp, P(scratch, torch.uint8), phases, t0, t1, ctas,
                                                  tn._lib.c_void_p(stream.cuda_stream)), "bwd part")


def timed(fn, reps=8):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(sA)
    for _ in range(reps):
        fn()
    e1.record(sA)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def sequential():
    part(7, 0, tiles, 0, sA)


def pipelined(q, gd, gw, first_full=True):
    bounds = [2 * ((tiles * i // q) // 2) for i in range(q)] + [tiles]
    part(1, 0, 0, 0, sA)
    last = None
    for i in range(q):
        part(2, bounds[i], bounds[i + 1], 0 if (i == 0 and first_full) else gd, sA)
        ev = torch.cuda.Event()
        ev.record(sA)
        sB.wait_event(ev)
        part(4, bounds[i], bounds[i + 1], gw if i < q - 1 else 0, sB)
    last = torch.cuda.Event()
    last.record(sB)
    sA.wait_event(last)


torch.cuda.synchronize()
sequential(); torch.cuda.synchronize()
ref = [g.clone() for g in grads]
print(f"rows {m}, tiles {tiles}")
print(f"sequential backward (zero + dgrad + wgrad): {timed(sequential):7.0f} us")
print("kernel alone at reduced grids (us):")
for g in (0, 111, 74):
    print(f"  dgrad ctas {g or 148:3d}: {timed(lambda: part(2, 0, tiles, g, sA)):7.0f}", flush=True)
for g in (0, 111, 74, 37):
    print(f"  wgrad ctas {g or 148:3d}: {timed(lambda: part(4, 0, tiles, g, sA)):7.0f}", flush=True)
print("pipelined (parts, dgrad ctas, wgrad ctas): us, max rel err of the gradients vs sequential")
for q, gd, gw in ((2, 96, 52), (4, 88, 60), (8, 88, 60)):
    us = timed(lambda: pipelined(q, gd, gw))
    torch.cuda.synchronize()
    err = max(float((a - b).abs().max() / (b.abs().max() + 1e-20)) for a, b in zip(grads, ref))
    print(f"  q={q:2d} gd={gd:3d} gw={gw:3d}: {us:7.0f} us   err {err:.2e}", flush=True)

# ---- what would wgrad run at if a producer kept its operands in L2?  (debug wrap: tile t reads tile t % wrap)
print("wgrad with operands wrapped onto an L2-resident window of tiles (timing only):")
for wg, wx in ((0, 0), (24, 0), (0, 24), (24, 24), (8, 8)):
    lib.nerf_debug_set_wgrad_wrap(wg, wx)
    print(f"  wrap G {wg:3d}  wrap X {wx:3d}: {timed(lambda: part(4, 0, tiles, 0, sA)):7.0f} us", flush=True)
lib.nerf_debug_set_wgrad_wrap(0, 0)
