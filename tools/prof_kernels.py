"""The four tensor-core kernels timed alone on the fine pass's rows (CUDA events, 10 launches each after 3 warm-ups).
NERF_B200_LIB_SUFFIX selects an experimental build (see build.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_nerf_b200 as tn
lib = tn._lib.load()
P = tn._lib.ptr
n, s = 4096, 192
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
g_s = torch.randn(m, device="cuda") * 1e-3; g_c = torch.randn(m, 3, device="cuda") * 1e-3
grads = [torch.zeros_like(p) for p in net.ordered_parameters()]
gp = tn._lib.pointer_array(grads)
def fwd(c):
    tn._lib.check(lib.nerf_mlp_bf16_forward(P(packed, torch.uint8), None, None, P(ray_o), P(ray_d), P(t), s, m, P(sig), P(rgb),
                                            P(c, torch.uint8) if c is not None else None, tn._lib.stream()), "fwd")
def bwd(ph):
    tn._lib.check(lib.nerf_mlp_bf16_backward_part(P(packed, torch.uint8), P(cache, torch.uint8), P(rgb), m, P(g_s), P(g_c), gp,
                                                  P(scratch, torch.uint8), ph, 0, tiles, 0, tn._lib.stream()), "bwd")
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
fwd(cache); bwd(7); torch.cuda.synchronize()
import time
out = []
for name, fn in (("fwd_inference", lambda: fwd(None)), ("fwd_train", lambda: fwd(cache)), ("dgrad", lambda: bwd(2)), ("wgrad", lambda: bwd(4))):
    time.sleep(1.0)
    out.append(f"{name} {timed(fn):7.1f} us")
print(os.environ.get("NERF_B200_LIB_SUFFIX", "(product)"), " | ".join(out))
