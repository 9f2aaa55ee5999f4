"""Render-pass orchestration.  Mirror of `VolumeRenderer` (reference src/renderer/volume_renderer.py:15-289):
same constructor, `camera` setter, `render_scene(...)` arguments / returns / ValueErrors.  Differences are
internal: rays are generated on the GPU from flat pixel ids (no (H*W,2) int64 screen-coordinate table is built
unless `.screen_coords` is read), and all per-ray work runs in libnerf_b200 kernels."""
from __future__ import annotations

from typing import Optional, Tuple, Union

import numpy as np
import torch

from .cameras import PerspectiveCamera
from .integrators import IntegratorBase
from .ray_samplers import RaySamplerBase, _cuda_device


class VolumeRenderer:
    def __init__(self, integrator: IntegratorBase, sampler: RaySamplerBase, camera: Optional[PerspectiveCamera] = None):
        self._integrator = integrator
        self._sampler = sampler
        self._camera = camera
        self._screen_coords = None
        if not self._camera:
            print("Warning: Camera parameters are not initialized.")

    def render_scene(self, target_scene, num_pixels: int, num_samples: Union[int, Tuple[int, int]], project_to_ndc: bool,
                     device: int, pixel_indices: Optional[torch.Tensor] = None, weights: Optional[torch.Tensor] = None,
                     num_ray_batch: int = None, uniforms=None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Returns (pixel_rgb (N,3), pixel_to_render (N,), weights (N,S)) -- volume_renderer.py:59-169.
        `uniforms` (extension) replays given uniform draws instead of the torch generator."""
        if not isinstance(num_pixels, int):
            raise ValueError(f"Expected variable of type int. Got {type(num_pixels)}.")
        if isinstance(num_samples, (tuple, list)):
            if len(num_samples) != 2:
                raise ValueError(
                    "Expected a tuple of length 2 for num_samples of type tuple. "
                    f"Got a tuple of length {len(num_samples)}."
                )
            if pixel_indices is None:
                raise ValueError(
                    "Expected a predefined set of pixels to render in hierarchical sampling. "
                    "Pixel indices are not provided."
                )
        dev = _cuda_device(device)
        total = self.camera.img_height * self.camera.img_width
        # volume_renderer.py:118-133
        if pixel_indices is not None:
            pixel_to_render = pixel_indices
        elif num_pixels < total:
            pixel_to_render = torch.tensor(np.random.choice(total, size=[num_pixels], replace=False))
        else:
            pixel_to_render = torch.arange(0, total)
        ray_bundle = self.sampler.generate_rays_from_pixels(pixel_to_render, self.camera, project_to_ndc, device=dev)
        sample_pts, ray_dir, delta_t = self.sampler.sample_along_rays(
            ray_bundle, num_samples, device=dev, weights=weights, uniforms=uniforms
        )
        pixel_rgb, weights, _, _ = self._render_ray_batches(
            target_scene, sample_pts, ray_dir, delta_t, num_batch=1 if num_ray_batch is None else num_ray_batch
        )
        return pixel_rgb, pixel_to_render, weights

    def _generate_screen_coords(self) -> torch.Tensor:
        """(H*W, 2) int64 (u = col, v = H-1-row) -- volume_renderer.py:171-190."""
        h, w = self.camera.img_height, self.camera.img_width
        p = torch.arange(h * w)
        return torch.stack([p % w, (h - 1) - p // w], dim=-1)

    def _render_ray_batches(self, target_scene, sample_pts, ray_dir, delta_t, num_batch: int):
        """Chunked query + integration (volume_renderer.py:192-261); chunk bounds follow
        torch.linspace(0, N, num_batch+1, dtype=long)."""
        rgb, weights, sigma, radiance = [], [], [], []
        n = sample_pts.shape[0]
        partitions = torch.linspace(0, n, num_batch + 1, dtype=torch.long)
        partitions[-1] = n
        bounds = partitions.tolist()
        for start, end in zip(bounds[:-1], bounds[1:]):
            sigma_b, radiance_b = target_scene.query_points(sample_pts[start:end], ray_dir[start:end])
            rgb_b, weights_b = self.integrator.integrate_along_rays(sigma_b, radiance_b, delta_t[start:end])
            rgb.append(rgb_b)
            weights.append(weights_b)
            sigma.append(sigma_b)
            radiance.append(radiance_b)
        return torch.cat(rgb, dim=0), torch.cat(weights, dim=0), torch.cat(sigma, dim=0), torch.cat(radiance, dim=0)

    @property
    def camera(self) -> PerspectiveCamera:
        return self._camera

    @property
    def integrator(self) -> IntegratorBase:
        return self._integrator

    @property
    def sampler(self) -> RaySamplerBase:
        return self._sampler

    @property
    def screen_coords(self) -> torch.Tensor:
        assert self._camera is not None, "Screen coordinates must not be None at rendering time."
        if self._screen_coords is None:
            self._screen_coords = self._generate_screen_coords()
        return self._screen_coords

    @camera.setter
    def camera(self, new_camera: PerspectiveCamera) -> None:
        self._camera = new_camera
        self._screen_coords = None  # rebuilt lazily; the kernels derive (u, v) from the pixel id
