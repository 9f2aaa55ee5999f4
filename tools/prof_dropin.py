"""Host-side profile of one training iteration through the drop-in API (VolumeRenderer.render_scene + autograd + Adam)."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import torch_nerf_b200 as tn

dev = torch.device("cuda", 0)
IMG, SC, SF, n_rays = 800, 64, 128, 4096
intr, c2w = bench._ref_scene(IMG)
torch.manual_seed(0)
nets = [tn.NeRF(63, 27, precision="bf16").to(dev) for _ in range(2)]
enc = {"coord_enc": tn.PositionalEncoder(3, 10, True), "dir_enc": tn.PositionalEncoder(3, 4, True)}
scenes = [tn.PrimitiveCube(net, enc) for net in nets]
ren = tn.VolumeRenderer(tn.QuadratureIntegrator(), tn.StratifiedSampler(), tn.PerspectiveCamera(intr, c2w, 2.0, 6.0))
opt = torch.optim.Adam([p for net in nets for p in net.parameters()], lr=5e-4, eps=1e-8)
loss_fn = torch.nn.MSELoss()
gt = torch.rand(IMG * IMG, 3)

def step(pix=None):
    opt.zero_grad()
    ren.camera = tn.PerspectiveCamera(intr, c2w, 2.0, 6.0)
    pred_c, idx, w_c = ren.render_scene(scenes[0], n_rays, SC, False, 0, pixel_indices=pix)
    loss = loss_fn(gt[idx].to(dev), pred_c)
    pred_f, idx_f, _ = ren.render_scene(scenes[1], n_rays, (SC, SF), False, 0, pixel_indices=idx, weights=w_c)
    loss = loss + loss_fn(gt[idx_f].to(dev), pred_f)
    loss.backward()
    opt.step()

for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): step()
t_host = (time.perf_counter() - t0) / 10
torch.cuda.synchronize()
t_all = (time.perf_counter() - t0) / 10
t0 = time.perf_counter(); [np.random.choice(IMG * IMG, size=[n_rays], replace=False) for _ in range(10)]; t_choice = (time.perf_counter() - t0) / 10
pix = torch.randperm(IMG * IMG)[:n_rays]
for _ in range(3): step(pix)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): step(pix)
t_host_p = (time.perf_counter() - t0) / 10
torch.cuda.synchronize()
t_all_p = (time.perf_counter() - t0) / 10
print(f"drop-in step: host enqueue {t_host*1e3:.2f} ms, wall {t_all*1e3:.2f} ms; np.random.choice alone {t_choice*1e3:.2f} ms; "
      f"with given pixel_indices: host {t_host_p*1e3:.2f} ms, wall {t_all_p*1e3:.2f} ms")
pr = cProfile.Profile(); pr.enable()
for _ in range(10): step(pix)
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
