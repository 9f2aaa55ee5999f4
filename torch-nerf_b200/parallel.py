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
