"""Diagnostic: per-step loss / gradient extremes of the 40-step fit, fp32 vs bf16."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch_nerf_b200 as tn
from torch_nerf_b200.engine import HotPathEngine
from oracle import nerf_oracle as orc

g = dict(np.load("tests/golden/train_step.npz"))
h, w, focal = int(g["h"]), int(g["w"]), float(g["focal"])
cam = tn.PerspectiveCamera({"f_x": focal, "f_y": focal, "img_width": w, "img_height": h}, torch.from_numpy(g["c2w"]), 2.0, 6.0)
gen = torch.Generator().manual_seed(5)
pix = torch.randperm(800 * 800, generator=gen)[:1024].cuda()
tgt = torch.stack([(pix % 800).float() / 800, (pix // 800).float() / 800, torch.full((1024,), 0.5, device="cuda")], -1).contiguous()
for seed in (11, 12):
    for precision in ("fp32", "bf16"):
        nets = []
        for s in (61, 62):
            net = tn.NeRF(63, 27, precision=precision)
            net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in orc.init_nerf_params(seed=s).items()})
            nets.append(net.cuda())
        eng = HotPathEngine(nets[0], nets[1], 64, 128, precision=precision)
        eng.enable_flat_params()
        opt = torch.optim.Adam([p for n_ in nets for p in n_.ordered_parameters()], lr=5e-4, eps=1e-8)
        torch.manual_seed(seed)
        print(f"=== {precision} seed {seed}")
        for it in range(40):
            losses = eng.train_pixels(cam, pix, tgt, False)
            gs_c = eng._get("bcgs", (1024, 64)); gs_f = eng._get("bfgs", (1024, 192))
            sig_f = eng.last["fine"]["sigma"]
            row0 = nets[1].fc_8.weight.grad[0].abs().max().item()
            gn = eng.flat.grad.norm().item()
            print(f"{it:2d} loss {losses.sum().item():.4f} max|g_sigma| c {gs_c.abs().max().item():.3e} f {gs_f.abs().max().item():.3e} "
                  f"|grad| {gn:.3e} fine fc_8 row0 max {row0:.3e} last-sigma>0 frac {float((sig_f[:, -1] > 0).float().mean()):.3f}")
            opt.step()
