"""Adam over the flat parameter buffer as ONE kernel launch (csrc/optim.cu).

`FlatAdam` is a torch.optim.Optimizer whose hyper-parameters, per-parameter state (`step`, `exp_avg`, `exp_avg_sq`)
and state_dict layout are torch.optim.Adam's, so the learning-rate schedulers and the checkpoint translation
(checkpoint.py) work on it unchanged; only the update itself goes through the C ABI.  torch's own fused Adam spends
~80 us on the single 1.19 M-element tensor (one multi-tensor chunk per 64 K elements); this launch takes ~8 us."""
from __future__ import annotations

import torch

from . import _lib


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, maximize=False, foreach=None,
                        capturable=False, differentiable=False, fused=None, decoupled_weight_decay=False)
        super().__init__(params, defaults)
        self.grad_scale = 1.0  # set to 1/world to fold the data-parallel average into the update
        self.exchange = None   # parallel.PeerExchange: step() then sums the ranks' gradients AND updates in one launch
        self._lib = _lib.load()

    def zero_grad(self, set_to_none: bool = True):
        """The kernels write gradients into the flat buffer by raw pointer (engine.FlatParams) and OVERWRITE it every
        iteration, so the reference loop's `optimizer.zero_grad()` (train.py:131; set_to_none by default) must never
        detach that buffer from the parameter: the gradients are zeroed in place and stay attached."""
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is not None:
                    p.grad.detach_()
                    p.grad.zero_()

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise ValueError("FlatAdam does not take a closure")
        for group in self.param_groups:
            if group.get("weight_decay", 0) != 0 or group.get("amsgrad", False) or group.get("maximize", False):
                raise ValueError("FlatAdam implements the reference's configuration only (no weight decay / amsgrad / maximize)")
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    raise RuntimeError("FlatAdam: a parameter has no gradient buffer attached (the kernels write into "
                                       "FlatParams.grad; do not set .grad to None)")
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous()):
                    raise ValueError("FlatAdam needs contiguous float32 CUDA parameters")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                if self.exchange is not None:
                    self.exchange.step(p, st["exp_avg"], st["exp_avg_sq"], float(group["lr"]), b1, b2, group["eps"],
                                       int(st["step"]), float(self.grad_scale))
                    continue
                with torch.cuda.device(p.device):
                    _lib.check(self._lib.nerf_adam_step(_lib.ptr(p), _lib.ptr(p.grad), _lib.ptr(st["exp_avg"]),
                                                        _lib.ptr(st["exp_avg_sq"]), p.numel(), float(group["lr"]), b1, b2,
                                                        group["eps"], int(st["step"]), float(self.grad_scale), _lib.stream()),
                               "nerf_adam_step")
        return None
