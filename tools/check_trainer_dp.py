"""N-rank run of torch_nerf_b200.Trainer (ray-sharded data parallel, fused peer exchange + Adam): ranks are seeded DIFFERENTLY on
purpose (the constructor must broadcast rank 0's replica), a few epochs are trained, replicas must stay bit-identical, a
checkpoint written by rank 0 must resume identically on all ranks, and render_image must all-gather a full frame.
   torchrun --nproc-per-node N --master-addr 127.0.0.1 --master-port 29521 tools/check_trainer_dp.py"""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch_nerf_b200 as tn
from oracle import nerf_oracle as orc
from torch_nerf_b200.parallel import init_distributed
from torch_nerf_b200.trainer import Trainer

rank, local, world = init_distributed("nccl")
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
torch.manual_seed(1000 + rank)  # different initial weights per rank: Trainer must fix that
c, f = tn.NeRF(63, 27, precision="bf16").to(dev), tn.NeRF(63, 27, precision="bf16").to(dev)
tr = Trainer(c, f, num_pixels=512, num_iter=1000, seed=3, rank=rank, world=world)
h = w = 96
focal = orc.blender_focal(w)
intr = {"f_x": focal, "f_y": focal, "img_width": w, "img_height": h}
gt = torch.full((h, w, 3), 0.5); gt[:, : w // 2, 0] = 0.8
views = [(gt, torch.from_numpy(orc.pose_spherical(40.0 * i, -30.0, 4.0))) for i in range(4)]

def checksum(t):
    s = t.detach().double().sum().reshape(1)
    out = [torch.zeros_like(s) for _ in range(world)]
    dist.all_gather(out, s)
    return [float(x) for x in out]

cs0 = checksum(tr.flat.flat)
first = tr.train_one_epoch(views, intr, epoch=0)
for ep in range(1, 4):
    last = tr.train_one_epoch(views, intr, epoch=10 + ep)
cs1 = checksum(tr.flat.flat)
st = tr.optimizer.state[tr.flat.param]
cs2 = checksum(st["exp_avg"])
img = tr.render_image(tn.PerspectiveCamera(intr, views[0][1], 2.0, 6.0))
ok_img = tuple(img.shape) == (3, h, w) and bool(torch.isfinite(img).all())
d = tempfile.mkdtemp() if rank == 0 else None
box = [d]
dist.broadcast_object_list(box, src=0)
if rank == 0:
    tr.save_ckpt(box[0], 14)
dist.barrier()
torch.manual_seed(77 + rank)
c2, f2 = tn.NeRF(63, 27, precision="bf16").to(dev), tn.NeRF(63, 27, precision="bf16").to(dev)
tr2 = Trainer(c2, f2, num_pixels=512, num_iter=1000, seed=9, rank=rank, world=world)
ep = tr2.load_ckpt(box[0])
cs3 = checksum(tr2.flat.flat)
if rank == 0:
    print(f"world {world}: exchange = {'peer-fused' if tr.exchange is not None else 'nccl'}")
    print("initial replicas identical:", len(set(cs0)) == 1, "| after 16 iterations:", len(set(cs1)) == 1, "| Adam state:", len(set(cs2)) == 1)
    print("losses first/last epoch:", first, last)
    print("render_image ok:", ok_img, "| resumed epoch", ep, "replicas identical:", len(set(cs3)) == 1, "== trained:", cs3[0] == cs1[0])
    assert len(set(cs0)) == 1 and len(set(cs1)) == 1 and len(set(cs2)) == 1 and len(set(cs3)) == 1 and cs3[0] == cs1[0] and ok_img
    assert last["coarse_loss"] < first["coarse_loss"]
    print("OK")
dist.destroy_process_group()
