"""Renders one 800x800 frame with the bf16 engine (for ncu launch lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import torch_nerf_b200 as tn
from torch_nerf_b200.engine import HotPathEngine

c = tn.NeRF(63, 27, precision="bf16").cuda(); f = tn.NeRF(63, 27, precision="bf16").cuda()
e = HotPathEngine(c, f, 64, 128, "bf16")
focal = bench.blender_focal(800)
cam = tn.PerspectiveCamera({"f_x": focal, "f_y": focal, "img_width": 800, "img_height": 800}, bench.pose_spherical(30., -30., 4.), 2.0, 6.0)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    e.render_frame(cam)
torch.cuda.synchronize()
print("done")
