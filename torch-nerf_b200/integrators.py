"""Alpha-compositing integrator backed by `nerf_composite_fwd` / `nerf_composite_bwd`.  Mirror of `IntegratorBase`
/ `QuadratureIntegrator` (reference src/renderer/integrators/integrator_base.py:8-26,
quadrature_integrator.py:14-67); backward replaces autograd through the same expression."""
from __future__ import annotations

import torch

from . import _lib


class IntegratorBase:
    def __init__(self, *arg, **kwargs):
        pass

    def integrate_along_rays(self, sigma, radiance, delta):
        raise NotImplementedError()


def composite_forward(sigma, radiance, delta, t=None, want_depth=False):
    lib = _lib.load()
    n, s = sigma.shape
    dev = sigma.device
    rgb = torch.empty((n, 3), device=dev, dtype=torch.float32)
    w = torch.empty((n, s), device=dev, dtype=torch.float32)
    depth = torch.empty((n,), device=dev, dtype=torch.float32) if want_depth else None
    opacity = torch.empty((n,), device=dev, dtype=torch.float32) if want_depth else None
    with torch.cuda.device(dev):
        _lib.check(
            lib.nerf_composite_fwd(_lib.ptr(sigma), _lib.ptr(radiance), _lib.ptr(delta), _lib.ptr(t) if want_depth else None,
                                   n, s, _lib.ptr(rgb), _lib.ptr(w), _lib.ptr(depth), _lib.ptr(opacity), _lib.stream()),
            "nerf_composite_fwd",
        )
    return rgb, w, depth, opacity


def composite_backward(sigma, radiance, delta, g_rgb, g_w):
    lib = _lib.load()
    n, s = sigma.shape
    g_sigma = torch.empty_like(sigma)
    g_rad = torch.empty_like(radiance)
    with torch.cuda.device(sigma.device):
        _lib.check(
            lib.nerf_composite_bwd(_lib.ptr(sigma), _lib.ptr(radiance), _lib.ptr(delta), _lib.ptr(g_rgb), _lib.ptr(g_w), n, s,
                                   _lib.ptr(g_sigma), _lib.ptr(g_rad), _lib.stream()),
            "nerf_composite_bwd",
        )
    return g_sigma, g_rad


class _Composite(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sigma, radiance, delta):
        sigma_c = sigma.detach().to(torch.float32).contiguous()
        rad_c = radiance.detach().to(torch.float32).contiguous()
        delta_c = delta.detach().to(torch.float32).contiguous()
        rgb, w, _, _ = composite_forward(sigma_c, rad_c, delta_c)
        ctx.save_for_backward(sigma_c, rad_c, delta_c)
        return rgb, w

    @staticmethod
    def backward(ctx, g_rgb, g_w):
        sigma, rad, delta = ctx.saved_tensors
        g_rgb = g_rgb.contiguous() if g_rgb is not None else torch.zeros((sigma.shape[0], 3), device=sigma.device)
        g_w = g_w.contiguous() if g_w is not None else None
        g_sigma, g_rad = composite_backward(sigma, rad, delta, g_rgb, g_w)
        return g_sigma, g_rad, None


class QuadratureIntegrator(IntegratorBase):
    def integrate_along_rays(self, sigma: torch.Tensor, radiance: torch.Tensor, delta: torch.Tensor):
        """sigma (N,S), radiance (N,S,3), delta (N,S) -> rgb (N,3), w (N,S)."""
        if not sigma.is_cuda:
            raise RuntimeError("torch_nerf_b200 runs on CUDA tensors only (no CPU fallback)")
        return _Composite.apply(sigma, radiance, delta)

    def integrate_with_depth(self, sigma, radiance, delta, t):
        """Addition (no reference counterpart): also returns depth = sum w_i t_i and opacity = sum w_i."""
        return composite_forward(sigma.contiguous(), radiance.contiguous(), delta.contiguous(), t.contiguous(), True)
