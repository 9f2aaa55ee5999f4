"""Launches the HBM-stage kernels (ray generation, coarse/fine sampling, compositing fwd/bwd) at the 800x800 frame's ray
count, each timed alone with CUDA events; under ncu it is the capture target for these kernels.
usage: python tools/prof_stages.py [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import torch_nerf_b200 as tn

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device("cuda:0")
lib = tn._lib.load()
P, st, ck = tn._lib.ptr, tn._lib.stream, tn._lib.check
nr, SC, SF = 800 * 800, 64, 128
S = SC + SF
focal = bench.blender_focal(800)
cam = tn.PerspectiveCamera({"f_x": focal, "f_y": focal, "img_width": 800, "img_height": 800}, bench.pose_spherical(30., -30., 4.), 2.0, 6.0).pack(False)
ro = torch.empty(nr, 3, device=dev); rd = torch.empty(nr, 3, device=dev)
u_c = torch.rand(nr, SC, device=dev); u1 = torch.rand(nr, SF, device=dev); u2 = torch.rand(nr, SF, device=dev)
w_c = torch.rand(nr, SC, device=dev) ** 6
t_c = torch.empty(nr, SC, device=dev); d_c = torch.empty(nr, SC, device=dev)
t_f = torch.empty(nr, S, device=dev); d_f = torch.empty(nr, S, device=dev)
sig = torch.rand(nr, S, device=dev); rad = torch.rand(nr, S, 3, device=dev)
rgb_o = torch.empty(nr, 3, device=dev); w_o = torch.empty(nr, S, device=dev)
g_rgb = torch.rand(nr, 3, device=dev); g_sig = torch.empty(nr, S, device=dev); g_rad = torch.empty(nr, S, 3, device=dev)
runs = [
    ("raygen", nr * 24, lambda: ck(lib.nerf_generate_rays_from_pixels(None, 0, nr, cam, P(ro), P(rd), st()), "raygen")),
    ("sample_coarse", nr * SC * 12, lambda: ck(lib.nerf_sample_coarse(P(ro), P(rd), nr, SC, 2.0, 6.0, P(u_c), P(t_c), None, None, P(d_c), st()), "coarse")),
    ("sample_fine", nr * (SC * 8 + SF * 8 + S * 8), lambda: ck(lib.nerf_sample_fine(P(ro), P(rd), nr, SC, SF, 2.0, 6.0, P(w_c), P(u_c), P(u1), P(u2), None, P(t_f), None, None, P(d_f), st()), "fine")),
    ("composite_fwd", nr * S * 24 + nr * 12, lambda: ck(lib.nerf_composite_fwd(P(sig), P(rad), P(d_f), None, nr, S, P(rgb_o), P(w_o), None, None, st()), "comp")),
    ("composite_bwd", nr * S * 36 + nr * 12, lambda: ck(lib.nerf_composite_bwd(P(sig), P(rad), P(d_f), P(g_rgb), None, nr, S, P(g_sig), P(g_rad), st()), "compb")),
]
for name, nbytes, fn in runs:
    fn(); fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3 / reps
    print(f"{name:14s} {sec * 1e6:9.1f} us  {nbytes / sec / 1e9:8.1f} GB/s")
