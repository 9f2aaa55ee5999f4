"""The NeRF MLP as an nn.Module whose forward/backward run in libnerf_b200.

Mirror of `NeRF` (reference src/network/nerf.py:11-136): same constructor, same `forward(pos, view_dir)` on
ENCODED inputs, same ValueErrors, same state_dict keys (fc_in, fc_1..fc_9, fc_out .weight/.bias, fp32 (out,in)),
so reference checkpoints load and Adam consumes `.parameters()` unchanged.

precision = "fp32": CUDA-core SGEMM chain (validation mode, <= 1e-3 parity gate)
precision = "bf16": tcgen05 tensor-core chain; needs RAW points/directions (the encoding is fused), so it is
                    reached through `query_raw` / PrimitiveCube.query_points, not through `forward`."""
from __future__ import annotations

import math
from typing import Tuple

import torch
import torch.nn as nn

from . import _lib

LAYER_NAMES = ("fc_in", "fc_1", "fc_2", "fc_3", "fc_4", "fc_5", "fc_6", "fc_7", "fc_8", "fc_9", "fc_out")


class _MlpF32(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dims, pos, view, *params):
        lib = _lib.load()
        m = pos.shape[0]
        dev = pos.device
        pos_c = pos.detach().to(torch.float32).contiguous()
        view_c = view.detach().to(torch.float32).contiguous()
        params_c = [p.detach() for p in params]
        sigma = torch.empty((m,), device=dev, dtype=torch.float32)
        rgb = torch.empty((m, 3), device=dev, dtype=torch.float32)
        cache = torch.empty((lib.nerf_mlp_f32_cache_floats(dims, m),), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(
                lib.nerf_mlp_f32_forward(dims, _lib.pointer_array(params_c), _lib.ptr(pos_c), _lib.ptr(view_c), m,
                                         _lib.ptr(sigma), _lib.ptr(rgb), _lib.ptr(cache), _lib.stream()),
                "nerf_mlp_f32_forward",
            )
        ctx.dims = dims
        ctx.save_for_backward(cache, rgb, *params_c)
        return sigma, rgb

    @staticmethod
    def backward(ctx, g_sigma, g_rgb):
        lib = _lib.load()
        cache, rgb, *params = ctx.saved_tensors
        m = rgb.shape[0]
        dev = rgb.device
        g_sigma = torch.zeros((m,), device=dev) if g_sigma is None else g_sigma.to(torch.float32).contiguous()
        g_rgb = torch.zeros((m, 3), device=dev) if g_rgb is None else g_rgb.to(torch.float32).contiguous()
        grads = [torch.empty_like(p) for p in params]
        scratch = torch.empty((lib.nerf_mlp_f32_bwd_scratch_floats(ctx.dims, m),), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(
                lib.nerf_mlp_f32_backward(ctx.dims, _lib.pointer_array(params), _lib.ptr(cache), _lib.ptr(rgb), m,
                                          _lib.ptr(g_sigma), _lib.ptr(g_rgb), _lib.pointer_array(grads),
                                          _lib.ptr(scratch), _lib.stream()),
                "nerf_mlp_f32_backward",
            )
        return (None, None, None, *grads)


class _MlpBf16(torch.autograd.Function):
    """Fused encoding + MLP on the tensor cores, training form: forward fills the activation cache (tile images +
    ReLU masks), backward runs the dgrad chain, wgrad and the head kernel."""

    @staticmethod
    def forward(ctx, net, pts, dirs, *params):
        lib = _lib.load()
        m = pts.shape[0]
        dev = pts.device
        pts_c = pts.detach().to(torch.float32).contiguous()
        dirs_c = dirs.detach().to(torch.float32).contiguous()
        packed = net.packed_weights()
        sigma = torch.empty((m,), device=dev, dtype=torch.float32)
        rgb = torch.empty((m, 3), device=dev, dtype=torch.float32)
        cache = torch.empty((lib.nerf_mlp_bf16_cache_bytes(m),), device=dev, dtype=torch.uint8)
        with torch.cuda.device(dev):
            _lib.check(
                lib.nerf_mlp_bf16_forward(_lib.ptr(packed, torch.uint8), _lib.ptr(pts_c), _lib.ptr(dirs_c), None, None, None,
                                          0, m, _lib.ptr(sigma), _lib.ptr(rgb), _lib.ptr(cache, torch.uint8), _lib.stream()),
                "nerf_mlp_bf16_forward",
            )
        ctx.net = net
        ctx.save_for_backward(cache, rgb, packed)
        ctx.shapes = [p.shape for p in params]
        return sigma, rgb

    @staticmethod
    def backward(ctx, g_sigma, g_rgb):
        lib = _lib.load()
        cache, rgb, packed = ctx.saved_tensors
        m = rgb.shape[0]
        dev = rgb.device
        g_sigma = torch.zeros((m,), device=dev) if g_sigma is None else g_sigma.to(torch.float32).contiguous()
        g_rgb = torch.zeros((m, 3), device=dev) if g_rgb is None else g_rgb.to(torch.float32).contiguous()
        grads = [torch.empty(s, device=dev, dtype=torch.float32) for s in ctx.shapes]
        scratch = torch.empty((lib.nerf_mlp_bf16_bwd_scratch_bytes(m),), device=dev, dtype=torch.uint8)
        with torch.cuda.device(dev):
            _lib.check(
                lib.nerf_mlp_bf16_backward(_lib.ptr(packed, torch.uint8), _lib.ptr(cache, torch.uint8), _lib.ptr(rgb), m,
                                           _lib.ptr(g_sigma), _lib.ptr(g_rgb), _lib.pointer_array(grads),
                                           _lib.ptr(scratch, torch.uint8), _lib.stream()),
                "nerf_mlp_bf16_backward",
            )
        return (None, None, None, *grads)


class NeRF(nn.Module):
    def __init__(self, pos_dim: int, view_dir_dim: int, feat_dim: int = 256, precision: str = "fp32"):
        super().__init__()
        self._pos_dim, self._view_dir_dim, self._feat_dim = pos_dim, view_dir_dim, feat_dim
        self.precision = precision
        f = feat_dim
        shapes = [(f, pos_dim)] + [(f, f)] * 4 + [(f, f + pos_dim)] + [(f, f)] * 2
        shapes += [(f + 1, f), (f // 2, f + view_dir_dim), (3, f // 2)]
        for name, (o, i) in zip(LAYER_NAMES, shapes):
            layer = nn.Module()
            # nn.Linear's default init (kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(in), 1/sqrt(in)) for both tensors)
            bound = 1.0 / math.sqrt(i)
            layer.weight = nn.Parameter(torch.empty(o, i).uniform_(-bound, bound))
            layer.bias = nn.Parameter(torch.empty(o).uniform_(-bound, bound))
            setattr(self, name, layer)
        self._dims = _lib.MlpDims(pos_dim, view_dir_dim, feat_dim)
        self._packed = None
        self._packed_version = None

    # ---- reference surface ------------------------------------------------------------------------
    def ordered_parameters(self):
        out = []
        for name in LAYER_NAMES:
            layer = getattr(self, name)
            out += [layer.weight, layer.bias]
        return out

    def forward(self, pos: torch.Tensor, view_dir: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """pos (M,pos_dim), view_dir (M,view_dir_dim) ENCODED -> sigma (M,), rgb (M,3)   (nerf.py:65-121)."""
        if (pos.ndim != 2) or (view_dir.ndim != 2):
            raise ValueError(f"Expected 2D tensors. Got {pos.ndim}, {view_dir.ndim}-D tensors.")
        if pos.shape[0] != view_dir.shape[0]:
            raise ValueError(f"The number of samples must match. Got {pos.shape[0]} and {view_dir.shape[0]}.")
        if pos.shape[-1] != self._pos_dim:
            raise ValueError(f"Expected {self._pos_dim}-D position vector. Got {pos.shape[-1]}.")
        if view_dir.shape[-1] != self._view_dir_dim:
            raise ValueError(f"Expected {self._view_dir_dim}-D view direction vector. Got {view_dir.shape[-1]}.")
        if not pos.is_cuda:
            raise RuntimeError("torch_nerf_b200 runs on CUDA tensors only (no CPU fallback)")
        # encoded inputs always take the fp32 chain: the tensor-core chain fuses the encoding (see query_raw)
        return _MlpF32.apply(self._dims, pos, view_dir, *self.ordered_parameters())

    @property
    def pos_dim(self) -> int:
        return self._pos_dim

    @property
    def view_dir_dim(self) -> int:
        return self._view_dir_dim

    @property
    def feat_dim(self) -> int:
        return self._feat_dim

    # ---- tensor-core path ---------------------------------------------------------------------------
    def supports_bf16(self) -> bool:
        return (self._pos_dim, self._view_dir_dim, self._feat_dim) == (63, 27, 256)

    def packed_weights(self, force: bool = False) -> torch.Tensor:
        """bf16 swizzled weight image for the tcgen05 chains, re-packed whenever a parameter changed (tracked through
        the parameters' version counters) or when `force` is set (e.g. after an optimizer stepped a flat buffer the
        parameters alias)."""
        lib = _lib.load()
        params = self.ordered_parameters()
        version = tuple(p._version for p in params) + tuple(p.data_ptr() for p in params)
        if force or self._packed is None or self._packed_version != version:
            dev = params[0].device
            if self._packed is None or self._packed.device != dev:
                self._packed = torch.empty((lib.nerf_mlp_bf16_packed_bytes(),), device=dev, dtype=torch.uint8)
            with torch.cuda.device(dev):
                _lib.check(
                    lib.nerf_mlp_bf16_pack(_lib.pointer_array([p.detach() for p in params]),
                                           _lib.ptr(self._packed, torch.uint8), _lib.stream()),
                    "nerf_mlp_bf16_pack",
                )
            self._packed_version = version
        return self._packed

    def query_raw(self, pts: torch.Tensor, dirs: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """The tcgen05 chain on RAW (M,3) points / directions (encoding fused in-kernel).  With autograd enabled
        the training form runs (activation cache + tensor-core backward)."""
        lib = _lib.load()
        if not self.supports_bf16():
            raise ValueError("the bf16 tensor-core chain is built for NeRF(63, 27, 256)")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return _MlpBf16.apply(self, pts, dirs, *self.ordered_parameters())
        m = pts.shape[0]
        dev = pts.device
        packed = self.packed_weights()
        sigma = torch.empty((m,), device=dev, dtype=torch.float32)
        rgb = torch.empty((m, 3), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(
                lib.nerf_mlp_bf16_forward(_lib.ptr(packed, torch.uint8), _lib.ptr(pts), _lib.ptr(dirs), None, None, None, 0,
                                          m, _lib.ptr(sigma), _lib.ptr(rgb), None, _lib.stream()),
                "nerf_mlp_bf16_forward",
            )
        return sigma, rgb
