"""Runs a few bf16 training steps of the fused engine at 4096 rays (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import torch_nerf_b200 as tn
from torch_nerf_b200.engine import HotPathEngine

torch.manual_seed(0)
c = tn.NeRF(63, 27, precision="bf16").cuda(); f = tn.NeRF(63, 27, precision="bf16").cuda()
eng = HotPathEngine(c, f, 64, 128, "bf16")
eng.enable_flat_params()
focal = bench.blender_focal(800)
cam = tn.PerspectiveCamera({"f_x": focal, "f_y": focal, "img_width": 800, "img_height": 800}, bench.pose_spherical(30., -30., 4.), 2.0, 6.0)
pix = torch.randperm(800 * 800)[:4096].cuda()
tgt = torch.rand(4096, 3).cuda()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    eng.train_pixels(cam, pix, tgt, False)
torch.cuda.synchronize()
print("ok")
