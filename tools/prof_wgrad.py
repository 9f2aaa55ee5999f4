"""Per-CTA timeline of the wgrad kernel (load balance of the static (unit, tile) partition)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_nerf_b200 as tn
lib = tn._lib.load()
P, VP = tn._lib.ptr, tn._lib.c_void_p
n, s = 4096, 192
m = n * s
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
def bwd(mask=4):
    tn._lib.check(lib.nerf_mlp_bf16_backward_part(P(packed, torch.uint8), P(cache, torch.uint8), P(rgb), m, P(g_sigma), P(g_rgb), gp,
                                                  P(scratch, torch.uint8), mask, 0, (m + 127) // 128, 0, tn._lib.stream()), "bwd")
bwd(7); torch.cuda.synchronize()
prof = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
lib.nerf_debug_set_wgrad_profile(VP(prof.data_ptr()))
bwd(); torch.cuda.synchronize()
lib.nerf_debug_set_wgrad_profile(None)
p = prof.cpu().view(148, 16)
t0 = int(p[:, 0].min())
print("cta unit nseg tiles0 tiles1 | start_us  acc_done_us  end_us | per stage of segment 0 (cycles): loader wait/total, mma wait/total, cuda wait/total")
for c in range(148):
    print(f"{c:3d} {int(p[c,3]):4d} {int(p[c,4]):4d} {int(p[c,5]):6d} {int(p[c,6]):6d} | {(int(p[c,0])-t0)/1e3:8.1f} {(int(p[c,1])-t0)/1e3:10.1f} {(int(p[c,2])-t0)/1e3:8.1f} | " + " ".join(f"{int(p[c,k])/max(1,int(p[c,14])):7.0f}" for k in (8, 9, 10, 11, 12, 13)))
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(5): bwd()
ev1.record(); torch.cuda.synchronize()
print(f"wgrad alone: {ev0.elapsed_time(ev1)/5*1e3:.0f} us")

