"""N-rank check + timing of the fused peer-memory exchange (csrc/dp_exchange.cu) against NCCL all-reduce + nerf_adam_step.
   torchrun --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/check_exchange.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch_nerf_b200 as tn
from torch_nerf_b200.parallel import PeerExchange, init_distributed
from torch_nerf_b200.optim import FlatAdam

rank, local, world = init_distributed("nccl")
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
n = 2 * 595844
px = PeerExchange(n, dev)
if rank == 0:
    print(f"world {world}: symmetric memory ok, multicast support: {getattr(px._h_grad, 'has_multicast_support', lambda *a: 'n/a')}")
torch.manual_seed(0)
p_a = torch.nn.Parameter(torch.randn(n, device=dev)); p_b = torch.nn.Parameter(p_a.detach().clone())
p_a.grad = px.grad
p_b.grad = torch.zeros(n, device=dev)
opt_a = FlatAdam([p_a], lr=5e-4, eps=1e-8); opt_a.grad_scale = 1.0 / world; opt_a.exchange = px
opt_b = FlatAdam([p_b], lr=5e-4, eps=1e-8); opt_b.grad_scale = 1.0 / world
g = torch.Generator(device=dev).manual_seed(100 + rank)
worst = 0.0
for it in range(5):
    gr = torch.randn(n, device=dev, generator=g) * (10.0 ** (it - 2))
    px.grad.copy_(gr); p_b.grad.copy_(gr)
    opt_a.step()
    dist.all_reduce(p_b.grad); opt_b.step()
    torch.cuda.synchronize()
    d_g = float((px.grad - p_b.grad).abs().max() / p_b.grad.abs().max())
    d_p = float((p_a - p_b).abs().max())
    worst = max(worst, d_g, d_p)
    if rank == 0:
        print(f"iter {it}: summed gradient rel diff vs NCCL {d_g:.2e}, parameter diff vs NCCL+Adam {d_p:.2e}")
# replicas identical?
chk = p_a.detach().double().sum().reshape(1).clone()
allc = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(allc, chk)
same = all(float(c) == float(allc[0]) for c in allc)
def timed(fn, reps=200):
    for _ in range(10): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
def nccl_path():
    dist.all_reduce(p_b.grad); opt_b.step()
t_fused = timed(opt_a.step)
t_nccl = timed(nccl_path)
if rank == 0:
    print(f"replicas bit-identical: {same}; worst diff {worst:.2e}")
    print(f"fused exchange+Adam: {t_fused:.1f} us/step   NCCL all-reduce + Adam: {t_nccl:.1f} us/step   (max over {world} ranks)")
    assert same and worst < 1e-5
dist.destroy_process_group()
