"""Pinhole camera record consumed by the ray-generation kernel.

Mirror of `PerspectiveCamera` (reference src/renderer/cameras.py:10-193): same constructor, same read-only
properties, same ValueError behaviour.  Besides the 4x4 intrinsic tensor it caches the packed
`nerf_camera_t` the C ABI takes."""
from __future__ import annotations

from typing import Dict, Tuple, Union

import torch

from . import _lib


class PerspectiveCamera:
    def __init__(self, intrinsic: Union[torch.Tensor, Dict[str, float]], extrinsic: torch.Tensor, t_near: float,
                 t_far: float):
        if not isinstance(intrinsic, (torch.Tensor, dict)):
            raise ValueError(f"Expected torch.Tensor of Python Dict as a camera intrinsic. Got {type(intrinsic)}.")
        self._extrinsic = extrinsic
        self._t_near = t_near
        self._t_far = t_far
        if isinstance(intrinsic, torch.Tensor):
            if intrinsic.shape != torch.Size((4, 4)):
                raise ValueError(f"Expected a tensor of shape (4, 4). Got {intrinsic.shape}.")
            self._intrinsic = intrinsic
            self._focal_x = float(intrinsic[0, 0])
            self._focal_y = float(intrinsic[1, 1])
            self._img_width = int(2 * intrinsic[0, 2])
            self._img_height = int(2 * intrinsic[1, 2])
        else:
            fx, fy = float(intrinsic["f_x"]), float(intrinsic["f_y"])
            w, h = float(intrinsic["img_width"]), float(intrinsic["img_height"])
            # cameras.py:109-117: principal point at the image centre, two dummy rows
            self._intrinsic = torch.tensor(
                [[fx, 0.0, w / 2.0, 0.0], [0.0, fy, h / 2.0, 0.0], [0.0, 0.0, 0.0, 0.0], [0.0, 0.0, -1.0, 0.0]],
                dtype=torch.float32,
            )
            self._focal_x, self._focal_y = fx, fy
            self._img_width, self._img_height = int(w), int(h)

    # ---- reference properties -------------------------------------------------------------------
    @property
    def intrinsic(self) -> torch.Tensor:
        return self._intrinsic

    @property
    def extrinsic(self) -> torch.Tensor:
        return self._extrinsic

    @property
    def t_near(self) -> float:
        return self._t_near

    @property
    def t_far(self) -> float:
        return self._t_far

    @property
    def img_width(self) -> int:
        return self._img_width

    @property
    def img_height(self) -> int:
        return self._img_height

    @property
    def focal_lengths(self) -> Tuple[float, float]:
        return (self._focal_x, self._focal_y)

    @intrinsic.setter
    def intrinsic(self, new_intrinsic: torch.Tensor) -> None:
        if not isinstance(new_intrinsic, torch.Tensor):
            raise ValueError(f"Expected variable of type torch.Tensor. Got {type(new_intrinsic)}.")
        if new_intrinsic.shape != torch.Size((4, 4)):
            raise ValueError(f"Expected tensor of shape (4, 4). Got {new_intrinsic.shape}.")
        self._intrinsic = new_intrinsic

    @extrinsic.setter
    def extrinsic(self, new_extrinsic: torch.Tensor) -> None:
        if not isinstance(new_extrinsic, torch.Tensor):
            raise ValueError(f"Expected variable of type torch.Tensor. Got {type(new_extrinsic)}.")
        if new_extrinsic.shape != torch.Size((4, 4)):
            raise ValueError(f"Expected tensor of shape (4, 4). Got {new_extrinsic.shape}.")
        self._extrinsic = new_extrinsic

    # ---- C-ABI view ------------------------------------------------------------------------------
    def pack(self, project_to_ndc: bool) -> "_lib.CameraStruct":
        return pack_camera(self, project_to_ndc)


def pack_camera(camera, project_to_ndc: bool) -> "_lib.CameraStruct":
    """Packs a camera for `nerf_generate_rays*` (include/nerf_b200.h nerf_camera_t).  `camera` is anything with the
    reference camera's read-only surface (cameras.py:120-153: intrinsic, extrinsic, t_near, img_width, img_height,
    focal_lengths) -- this package's PerspectiveCamera or the reference's own, so the sampler also works behind the
    reference's VolumeRenderer.  The NDC scale factors are evaluated in Python floats first (sampler_base.py:236-253
    multiplies float32 tensors by Python scalars, which torch applies as float32 scalars)."""
    k = camera.intrinsic.detach().to("cpu", torch.float32)
    e = camera.extrinsic.detach().to("cpu", torch.float32)
    cam = _lib.CameraStruct()
    cam.fx, cam.fy, cam.cx, cam.cy = float(k[0, 0]), float(k[1, 1]), float(k[0, 2]), float(k[1, 2])
    rot = e[:3, :3].reshape(-1).tolist()
    for i in range(9):
        cam.rot[i] = rot[i]
    trans = e[:3, -1].tolist()
    for i in range(3):
        cam.trans[i] = trans[i]
    cam.img_w, cam.img_h = int(camera.img_width), int(camera.img_height)
    cam.project_to_ndc = 1 if project_to_ndc else 0
    if project_to_ndc:
        fx, fy = camera.focal_lengths
        if fx != fy:
            raise ValueError(
                "Focal length used for computing NDC is ambiguous."
                f"Two different focal lengths ({fx}, {fy}) exists but only one can be used."
            )
        if camera.t_near < 0:
            raise ValueError(f"Expected a real number greater than or equal to 0. Got {camera.t_near}.")
        cam.ndc_sx = -(2 * fx / cam.img_w)
        cam.ndc_sy = -(2 * fx / cam.img_h)
        cam.ndc_two_near = 2 * camera.t_near
    return cam
