"""Runs the tensor-core forward chain on the fine pass's rows (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_nerf_b200 as tn

lib = tn._lib.load()
n, s = 4096, 192
m = n * s
net = tn.NeRF(63, 27, precision="bf16").cuda()
packed = net.packed_weights()
ray_o = torch.randn(n, 3, device="cuda")
ray_d = torch.randn(n, 3, device="cuda")
t = torch.rand(n, s, device="cuda") * 4 + 2
sig = torch.empty(m, device="cuda")
rgb = torch.empty(m, 3, device="cuda")
P = tn._lib.ptr
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    tn._lib.check(lib.nerf_mlp_bf16_forward(P(packed, torch.uint8), None, None, P(ray_o), P(ray_d), P(t), s, m, P(sig), P(rgb), None, tn._lib.stream()), "fwd")
torch.cuda.synchronize()
print("ok", float(sig.mean()), float(rgb.mean()))
