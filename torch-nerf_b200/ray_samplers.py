"""Ray generation and sampling along rays, on the GPU through libnerf_b200.

Mirrors of `RayBundle`, `RaySamplerBase.generate_rays` (reference src/renderer/ray_samplers/sampler_base.py:11-59,
134-197), `StratifiedSampler.sample_along_rays` (stratified_sampler.py:17-128) and `sample_pdf` (utils.py:8-58).
Uniform random numbers are drawn with torch on the device in the reference's order and shapes, so a run seeded
like the reference on torch-CUDA consumes the generator identically."""
from __future__ import annotations

from typing import Tuple, Union

import torch

from . import _lib
from .cameras import PerspectiveCamera, pack_camera


class RayBundle:
    """sampler_base.py:11-59."""

    def __init__(self, ray_origin: torch.Tensor, ray_dir: torch.Tensor, t_near: float, t_far: float, is_ndc: bool):
        self._ray_origin, self._ray_dir = ray_origin, ray_dir
        self._t_near, self._t_far, self._is_ndc = t_near, t_far, is_ndc

    @property
    def ray_origin(self) -> torch.Tensor:
        return self._ray_origin

    @property
    def ray_dir(self) -> torch.Tensor:
        return self._ray_dir

    @property
    def t_near(self) -> float:
        return self._t_near

    @property
    def t_far(self) -> float:
        return self._t_far

    @property
    def is_ndc(self) -> bool:
        return self._is_ndc


def _cuda_device(device=None) -> torch.device:
    if isinstance(device, torch.device):
        dev = device
    elif isinstance(device, int):
        dev = torch.device("cuda", device)
    elif isinstance(device, str):
        dev = torch.device(device)
    else:
        dev = torch.device("cuda", torch.cuda.current_device())
    if dev.type != "cuda":
        raise RuntimeError("torch_nerf_b200 runs on CUDA devices only (no CPU fallback)")
    return dev


class _PinnedStager:
    """Host -> device copies of small index tensors without the stream synchronisation a pageable `tensor.to(device)`
    implies (the reference keeps pixel indices on the host, volume_renderer.py:118-142; a synchronising copy at the start
    of every pass stalls the launch queue).  A ring of pinned buffers, each guarded by the event of its last copy."""

    def __init__(self, slots: int = 4):
        self._slots = [None] * slots
        self._next = 0

    def to_device(self, t: torch.Tensor, dev: torch.device, dtype: torch.dtype) -> torch.Tensor:
        if t.is_cuda:
            return t.to(device=dev, dtype=dtype).contiguous()
        src = t.detach().to(dtype).contiguous()
        i = self._next
        self._next = (i + 1) % len(self._slots)
        slot = self._slots[i]
        if slot is None or slot[0].numel() < src.numel() or slot[0].dtype != dtype:
            buf = torch.empty((max(src.numel(), 1),), dtype=dtype, pin_memory=True)
            slot = [buf, None]
            self._slots[i] = slot
        elif slot[1] is not None:
            slot[1].synchronize()  # the copy that last used this buffer (several passes ago) has finished
        view = slot[0][: src.numel()].view(src.shape)
        view.copy_(src)
        with torch.cuda.device(dev):
            out = view.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        slot[1] = ev
        return out


_stager = _PinnedStager()


def _f32_cuda(t: torch.Tensor, dev: torch.device) -> torch.Tensor:
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


def make_bins(t_near: float, t_far: float, num_partitions: int) -> Tuple[torch.Tensor, float]:
    """stratified_sampler.py:130-164 (`_create_t_bins`) evaluated by the library's host helper."""
    lib = _lib.load()
    buf = (_lib.c_float * num_partitions)()
    step = _lib.c_float()
    _lib.check(lib.nerf_make_bins(float(t_near), float(t_far), int(num_partitions), buf, step), "nerf_make_bins")
    return torch.tensor(list(buf), dtype=torch.float32), (t_far - t_near) / num_partitions


class RaySamplerBase:
    def __init__(self):
        pass

    def generate_rays(self, pixel_coords: torch.Tensor, camera: PerspectiveCamera, project_to_ndc: bool) -> RayBundle:
        """sampler_base.py:134-197.  pixel_coords (N,2) integer screen coordinates (CPU or CUDA)."""
        lib = _lib.load()
        dev = _cuda_device(pixel_coords.device if pixel_coords.is_cuda else None)
        coords = _stager.to_device(pixel_coords, dev, torch.int64)
        n = coords.shape[0]
        cam = pack_camera(camera, project_to_ndc)
        ray_o = torch.empty((n, 3), device=dev, dtype=torch.float32)
        ray_d = torch.empty((n, 3), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(
                lib.nerf_generate_rays(_lib.ptr(coords, torch.int64), n, cam, _lib.ptr(ray_o), _lib.ptr(ray_d), _lib.stream()),
                "nerf_generate_rays",
            )
        return RayBundle(ray_o, ray_d, t_near=camera.t_near, t_far=camera.t_far, is_ndc=project_to_ndc)

    def generate_rays_from_pixels(self, pixel_indices, camera: PerspectiveCamera, project_to_ndc: bool, device=None,
                                  first_pixel: int = 0, count: int = 0) -> RayBundle:
        """Fused form: flat pixel ids p = row*W + col (volume_renderer.py:171-190 folded into the kernel);
        `pixel_indices=None` renders the contiguous range [first_pixel, first_pixel + count)."""
        lib = _lib.load()
        dev = _cuda_device(device)
        if pixel_indices is not None:
            pix = _stager.to_device(pixel_indices, dev, torch.int64)
            n = pix.shape[0]
        else:
            pix, n = None, int(count)
        cam = pack_camera(camera, project_to_ndc)
        ray_o = torch.empty((n, 3), device=dev, dtype=torch.float32)
        ray_d = torch.empty((n, 3), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(
                lib.nerf_generate_rays_from_pixels(_lib.ptr(pix, torch.int64), int(first_pixel), n, cam, _lib.ptr(ray_o),
                                                   _lib.ptr(ray_d), _lib.stream()),
                "nerf_generate_rays_from_pixels",
            )
        return RayBundle(ray_o, ray_d, t_near=camera.t_near, t_far=camera.t_far, is_ndc=project_to_ndc)

    def map_rays_to_ndc(self, focal_length: float, z_near: float, img_height: int, img_width: int, ray_origin: torch.Tensor,
                        ray_dir: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """sampler_base.py:199-257: world-frame rays (N,3) -> NDC (no shift of the origin to the near plane, like the
        reference).  Returns (projected_origin, projected_dir)."""
        if z_near < 0:
            raise ValueError(f"Expected a real number greater than or equal to 0. Got {z_near}.")
        lib = _lib.load()
        dev = _cuda_device(ray_origin.device if ray_origin.is_cuda else None)
        o, d = _f32_cuda(ray_origin, dev), _f32_cuda(ray_dir, dev)
        out_o, out_d = torch.empty_like(o), torch.empty_like(d)
        with torch.cuda.device(dev):
            _lib.check(lib.nerf_map_rays_to_ndc(_lib.ptr(o), _lib.ptr(d), o.shape[0], float(focal_length), float(z_near),
                                                int(img_height), int(img_width), _lib.ptr(out_o), _lib.ptr(out_d), _lib.stream()),
                       "nerf_map_rays_to_ndc")
        return out_o, out_d

    def sample_along_rays(self, *args, **kwargs):
        raise NotImplementedError()


def sample_pdf(bins: torch.Tensor, partition_size: float, weights: torch.Tensor, num_sample: int,
               uniforms: Tuple[torch.Tensor, torch.Tensor] = None, return_indices: bool = False):
    """utils.py:8-58.  `bins` (N,S) must be the stratified partition (rows identical, uniform spacing
    `partition_size`), which is the only way the reference calls it (stratified_sampler.py:80-85); the kernel
    re-derives the bins from (bins[0,0], partition_size).  `weights` is modified in place (+= 1e-5).
    `uniforms=(u1, u2)` replays given draws instead of torch.rand / torch.rand_like."""
    lib = _lib.load()
    if not weights.is_cuda:
        raise RuntimeError("torch_nerf_b200 runs on CUDA tensors only (no CPU fallback)")
    dev = weights.device
    n, sc = weights.shape
    w = weights.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        raise ValueError("weights must be a contiguous float32 tensor (it is updated in place)")
    t_near = float(bins[0, 0])
    t_far = t_near + partition_size * sc
    if uniforms is None:
        u1 = torch.rand((n, num_sample), device=dev)
        u2 = torch.rand((n, num_sample), device=dev)  # == rand_like(t_start)
    else:
        u1, u2 = (_f32_cuda(u, dev) for u in uniforms)
    t_fine = torch.empty((n, num_sample), device=dev, dtype=torch.float32)
    idx = torch.empty((n, num_sample), device=dev, dtype=torch.int64) if return_indices else None
    with torch.cuda.device(dev):
        _lib.check(
            lib.nerf_sample_pdf(t_near, t_far, _lib.ptr(w), _lib.ptr(u1), _lib.ptr(u2), n, sc, int(num_sample),
                                _lib.ptr(t_fine), _lib.ptr(idx, torch.int64), _lib.stream()),
            "nerf_sample_pdf",
        )
    return (t_fine, idx) if return_indices else t_fine


class StratifiedSampler(RaySamplerBase):
    """stratified_sampler.py:12-164."""

    def sample_along_rays(self, ray_bundle: RayBundle, num_samples: Union[int, Tuple[int, int]], device: int,
                          weights: torch.Tensor = None, uniforms=None, return_extras: bool = False,
                          materialize: bool = True):
        """Returns (sample_pts (N,S,3), ray_dir (N,S,3), delta (N,S)) like the reference.

        Extensions (keyword-only in spirit): `uniforms` replays given draws -- (u,) for the coarse branch,
        (u0, u1, u2) for the hierarchical one; `return_extras` appends a dict with `t` (N,S) and, for the
        hierarchical branch, the int64 bin indices `idx`; `materialize=False` skips the (N,S,3) tensors
        (returns None for them) for the fused renderer."""
        lib = _lib.load()
        dev = _cuda_device(device)
        ray_o = _f32_cuda(ray_bundle.ray_origin, dev)
        ray_d = _f32_cuda(ray_bundle.ray_dir, dev)
        n = ray_o.shape[0]
        near, far = float(ray_bundle.t_near), float(ray_bundle.t_far)
        extras = {}
        if weights is not None:  # hierarchical sampling (stratified_sampler.py:57-90)
            if not isinstance(weights, torch.Tensor):
                raise ValueError(f"Expected an instance of torch.Tensor. Got {type(weights)}.")
            if not isinstance(num_samples, (tuple, list)):
                raise ValueError(
                    "Expected a tuple for parameter 'num_samples' when hierarchical sampling is used. "
                    f"Got a parameter of type {type(num_samples)}."
                )
            sc, sf = num_samples
            s = sc + sf
            w = weights.detach()
            if not (w.is_cuda and w.device == dev and w.dtype == torch.float32 and w.is_contiguous()):
                # the reference moves the tensor with .to(); the in-place += 1e-5 then lands on that copy
                w = _f32_cuda(w, dev)
            if uniforms is None:
                u0 = torch.rand((n, sc), device=dev)   # rand_like(t_bins)            :77
                u1 = torch.rand((n, sf), device=dev)   # rand((N, Sf))                utils.py:43
                u2 = torch.rand((n, sf), device=dev)   # rand_like(t_start)           utils.py:56
            else:
                u0, u1, u2 = (_f32_cuda(u, dev) for u in uniforms)
            idx = torch.empty((n, sf), device=dev, dtype=torch.int64) if return_extras else None
        else:  # stratified_sampler.py:91-109
            if not isinstance(num_samples, int):
                raise ValueError(
                    "Expected an integer for parameter 'num_samples' when hierarchical sampling is unused. "
                    f"Got a parameter of type {type(num_samples)}."
                )
            s = num_samples
            u = torch.rand((n, s), device=dev) if uniforms is None else _f32_cuda(uniforms[0], dev)
        t = torch.empty((n, s), device=dev, dtype=torch.float32) if (return_extras or not materialize) else None
        pts = torch.empty((n, s, 3), device=dev, dtype=torch.float32) if materialize else None
        dirs = torch.empty((n, s, 3), device=dev, dtype=torch.float32) if materialize else None
        delta = torch.empty((n, s), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            if weights is not None:
                _lib.check(
                    lib.nerf_sample_fine(_lib.ptr(ray_o), _lib.ptr(ray_d), n, int(sc), int(sf), near, far, _lib.ptr(w),
                                         _lib.ptr(u0), _lib.ptr(u1), _lib.ptr(u2), _lib.ptr(idx, torch.int64),
                                         _lib.ptr(t), _lib.ptr(pts), _lib.ptr(dirs), _lib.ptr(delta), _lib.stream()),
                    "nerf_sample_fine",
                )
                extras["idx"] = idx
            else:
                _lib.check(
                    lib.nerf_sample_coarse(_lib.ptr(ray_o), _lib.ptr(ray_d), n, int(s), near, far, _lib.ptr(u),
                                           _lib.ptr(t), _lib.ptr(pts), _lib.ptr(dirs), _lib.ptr(delta), _lib.stream()),
                    "nerf_sample_coarse",
                )
        extras["t"] = t
        extras["ray_o"], extras["ray_d"] = ray_o, ray_d
        if return_extras:
            return pts, dirs, delta, extras
        return pts, dirs, delta
