"""Ray-sharded data parallelism (the reference is single-device, runner_utils.py:431-453).

One process per GPU, identical replicas of both networks.  Rays are independent, so the global batch is split by
ray index with no data-path collective; the only exchange per step is one all-reduce of the flat gradient buffer
(2 x 595 844 floats = 4.77 MB) followed by a 1/world scale.  With equal shard sizes the mean of the per-rank MSE
gradients equals the gradient of the global-batch MSE (runner_utils.py:731 nn.MSELoss = mean over all rays)."""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def init_distributed(backend: str | None = None) -> Tuple[int, int, int]:
    """Initialises torch.distributed from the torchrun environment.  Returns (rank, local_rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kwargs["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, rank=rank, world_size=world, **kwargs)
    return rank, local, world


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, end) slice of n units for `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_rays(pixel_indices: torch.Tensor, target: torch.Tensor, rank: int, world: int):
    """The slice of a global ray batch (pixel ids + target colours) this rank renders."""
    a, b = shard_range(pixel_indices.shape[0], rank, world)
    return pixel_indices[a:b], target[a:b]


def allreduce_mean_(flat_grad: torch.Tensor, world: int | None = None, scale: bool = True) -> torch.Tensor:
    """In-place average of the flat gradient buffer over all ranks (sum all-reduce, then scale).  With scale=False
    only the sum is formed: optim.FlatAdam folds the 1/world into its update (grad_scale)."""
    if not (dist.is_available() and dist.is_initialized()):
        return flat_grad
    world = dist.get_world_size() if world is None else world
    if world == 1:
        return flat_grad
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    if scale:
        flat_grad.mul_(1.0 / world)
    return flat_grad


def broadcast_replica_state(param: torch.Tensor, optimizer: torch.optim.Optimizer, scheduler=None, src: int = 0) -> None:
    """Makes every rank's replica identical to rank `src`'s: the flat parameter buffer, its Adam state (`step`,
    `exp_avg`, `exp_avg_sq`), the learning rates and the scheduler position.  Called by Trainer at construction and
    after load_ckpt, so replicas cannot silently diverge through per-rank seeds or per-rank checkpoint files."""
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("broadcast_replica_state needs an initialised torch.distributed process group")
    if dist.get_world_size() == 1:
        return
    dev = param.device
    dist.broadcast(param.data, src=src)
    st = optimizer.state.get(param, {})
    has = torch.tensor([1 if len(st) else 0], device=dev)
    dist.broadcast(has, src=src)
    if int(has.item()):
        if not len(st):  # this rank has no state yet: create it so the broadcast has somewhere to land
            st = optimizer.state[param]
            st["step"] = torch.tensor(0.0, dtype=torch.float32)
            st["exp_avg"] = torch.zeros_like(param.data)
            st["exp_avg_sq"] = torch.zeros_like(param.data)
        step = st["step"].detach().to(dev, torch.float32).reshape(1).clone()
        dist.broadcast(step, src=src)
        st["step"] = step.reshape(()).to(st["step"].device)
        dist.broadcast(st["exp_avg"], src=src)
        dist.broadcast(st["exp_avg_sq"], src=src)
    else:
        optimizer.state.pop(param, None)
    lrs = torch.tensor([float(g["lr"]) for g in optimizer.param_groups], device=dev, dtype=torch.float64)
    dist.broadcast(lrs, src=src)
    for g, lr in zip(optimizer.param_groups, lrs.tolist()):
        g["lr"] = lr
    if scheduler is not None:
        pos = torch.tensor([float(scheduler.last_epoch)], device=dev, dtype=torch.float64)
        dist.broadcast(pos, src=src)
        scheduler.last_epoch = int(pos.item())
        scheduler._last_lr = [g["lr"] for g in optimizer.param_groups]
