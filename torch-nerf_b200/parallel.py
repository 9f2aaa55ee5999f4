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


class PeerExchange:
    """The data-parallel step as ONE kernel over NVLink peer memory (csrc/dp_exchange.cu, nerf_dp_exchange_adam): the flat
    gradient buffer lives in symmetric memory every rank maps, each rank reduces its slice of all buffers with P2P loads,
    stores the sum back into all of them with P2P stores, and runs Adam -- instead of an NCCL all-reduce followed by an
    optimizer launch.  Construction is collective (every rank of `group`); it raises when symmetric memory is not
    available on the box, and callers then stay on `allreduce_mean_` + FlatAdam.

        px = PeerExchange(numel, device)          # px.grad: the flat gradient buffer the kernels must write into
        flat = engine.enable_flat_params(grad_buffer=px.grad)
        opt = FlatAdam([flat.param], ...); opt.exchange = px     # opt.step() now makes the fused call
    """

    def __init__(self, numel: int, device: torch.device, group=None):
        import ctypes

        import torch.distributed._symmetric_memory as symm

        from . import _lib

        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerExchange needs an initialised torch.distributed process group")
        group = dist.group.WORLD if group is None else group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 16:
            raise RuntimeError("PeerExchange supports up to 16 ranks of one NVLink domain")
        self.numel = int(numel)
        padded = (self.numel + 3) // 4 * 4
        try:  # older torch releases want the group registered first; newer ones deprecate the call (FutureWarning)
            import warnings

            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                symm.enable_symm_mem_for_group(group.group_name)
        except Exception:
            pass
        with torch.cuda.device(device):
            self._grad_full = symm.empty(padded, dtype=torch.float32, device=device)
            self._flags = symm.empty(64, dtype=torch.int32, device=device)
            self._grad_full.zero_()
            self._flags.zero_()
            self._h_grad = symm.rendezvous(self._grad_full, group.group_name)
            self._h_flags = symm.rendezvous(self._flags, group.group_name)
            self.counter = torch.zeros(1, device=device, dtype=torch.int32)
            torch.cuda.synchronize(device)
        dist.barrier(group)  # every pad is zero before anybody's first kernel writes a flag
        self.grad = self._grad_full[: self.numel]
        self._grad_ptrs = (ctypes.c_void_p * self.world)(*[int(p) for p in self._h_grad.buffer_ptrs])
        self._flag_ptrs = (ctypes.c_void_p * self.world)(*[int(p) for p in self._h_flags.buffer_ptrs])
        self.seq = 0
        self._lib = _lib.load()
        self.device = device

    def step(self, param: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, lr: float, beta1: float, beta2: float,
             eps: float, step: int, grad_scale: float) -> None:
        """Collective: sums the gradient buffers of all ranks and applies Adam update number `step` (1-based)."""
        from . import _lib

        if param.grad is None or param.grad.data_ptr() != self.grad.data_ptr():
            raise RuntimeError("PeerExchange: the parameter's gradient is not the symmetric buffer (use "
                               "engine.enable_flat_params(grad_buffer=px.grad))")
        self.seq += 1
        with torch.cuda.device(self.device):
            _lib.check(self._lib.nerf_dp_exchange_adam(self._grad_ptrs, self._flag_ptrs, self.rank, self.world, _lib.ptr(param),
                                                       _lib.ptr(exp_avg), _lib.ptr(exp_avg_sq), self.numel, float(lr), float(beta1),
                                                       float(beta2), float(eps), int(step), float(grad_scale), self.seq,
                                                       _lib.ptr(self.counter, torch.int32), _lib.stream()),
                       "nerf_dp_exchange_adam")


    def allreduce_sum_(self) -> torch.Tensor:
        """Collective: the exchange alone -- every rank's `grad` ends up holding the sum over the ranks (no update)."""
        from . import _lib

        self.seq += 1
        with torch.cuda.device(self.device):
            _lib.check(self._lib.nerf_dp_exchange_adam(self._grad_ptrs, self._flag_ptrs, self.rank, self.world, None, None, None,
                                                       self.numel, 0.0, 0.9, 0.999, 1e-8, 1, 1.0, self.seq,
                                                       _lib.ptr(self.counter, torch.int32), _lib.stream()),
                       "nerf_dp_exchange_adam")
        return self.grad


def make_exchange(numel: int, device: torch.device, world: int):
    """PeerExchange when the box supports it (world > 1, symmetric memory), else None -> NCCL all-reduce + FlatAdam."""
    if world <= 1 or os.environ.get("NERF_B200_EXCHANGE", "peer") == "nccl":
        return None
    ok = 1
    px = None
    try:
        px = PeerExchange(numel, device)
    except Exception as err:  # noqa: BLE001 -- any failure (no symmetric memory, no P2P) selects the NCCL path
        import warnings

        warnings.warn(f"PeerExchange unavailable, using NCCL all-reduce + Adam: {type(err).__name__}: {err}")
        ok = 0
    flag = torch.tensor([ok], device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)  # all ranks or none
    return px if int(flag.item()) == 1 else None
