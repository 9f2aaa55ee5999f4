"""GPU idle time inside a training step: kernel-time sum vs span (torch profiler, CUPTI)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import torch_nerf_b200 as tn
from torch_nerf_b200.engine import HotPathEngine
from torch.profiler import profile, ProfilerActivity

torch.manual_seed(0)
c = tn.NeRF(63, 27, precision="bf16").cuda(); f = tn.NeRF(63, 27, precision="bf16").cuda()
eng = HotPathEngine(c, f, 64, 128, "bf16")
flat = eng.enable_flat_params()
from torch_nerf_b200.optim import FlatAdam
opt = FlatAdam([flat.param], lr=5e-4)
focal = bench.blender_focal(800)
cam = tn.PerspectiveCamera({"f_x": focal, "f_y": focal, "img_width": 800, "img_height": 800}, bench.pose_spherical(30., -30., 4.), 2.0, 6.0)
pix = torch.randperm(800 * 800)[:4096].cuda()
tgt = torch.rand(4096, 3).cuda()
losses = torch.zeros(2, device="cuda")
def step():
    eng.train_pixels(cam, pix, tgt, False, loss_out=losses)
    opt.step()
for _ in range(5): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
busy = sum(e.time_range.end - e.time_range.start for e in evs)
print(f"span {t1 - t0:.0f} us for 5 steps, kernel time {busy:.0f} us, idle {(t1 - t0 - busy):.0f} us ({100 * (t1 - t0 - busy) / (t1 - t0):.1f}%)")
gaps = []
for a, b in zip(evs[:-1], evs[1:]):
    g = b.time_range.start - a.time_range.end
    if g > 0: gaps.append((g, a.name[:50], b.name[:50]))
gaps.sort(reverse=True)
for g in gaps[:25]: print(f"{g[0]:8.1f} us  after {g[1]:50s} before {g[2]}")
import collections
agg = collections.Counter()
for e in evs: agg[e.name[:60]] += e.time_range.end - e.time_range.start
for k, v in agg.most_common(12): print(f"{v / 5:9.1f} us/step  {k}")
