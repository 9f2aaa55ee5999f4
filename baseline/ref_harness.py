"""Drives the UNMODIFIED reference hot-path modules staged under baseline/_ref/ (see stage_reference.py) the way the
reference's own runners do, on CPU or on torch-CUDA.  Benchmark / test infrastructure only: nothing in the product
package imports this file, and nothing here is ever on the B200 arm's timed path.

    RefSession(device)            renderer + coarse/fine scenes + Adam + ExponentialLR + MSELoss, as the factories of
                                  runners/runner_utils.py:526-550, 569-660, 663-715, 718-733 build them
    RefSession.train_step(...)    the loop body of runners/train.py:130-218 (hydra cannot be imported, so the ~25 lines
                                  are restated; every call inside goes to reference code)
    RefSession.render_frame(...)  runners/render.py:58-107

The reference modules are imported under their own package name (`torch_nerf`) from baseline/_ref after the manifest of
file hashes has been re-checked, so "unmodified" is verifiable on the GPU box too.
"""
from __future__ import annotations

import hashlib
import importlib
import json
import os
import sys
from typing import Optional, Tuple

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")
_mods = None


def available() -> Tuple[bool, str]:
    man = os.path.join(REF_ROOT, "MANIFEST.json")
    if not os.path.exists(man):
        return False, "baseline/_ref is not staged (run baseline/stage_reference.py in the dev container)"
    return True, ""


def load():
    """Imports the staged reference modules (once) and returns them as a dict of the classes the hot path uses."""
    global _mods
    if _mods is not None:
        return _mods
    ok, why = available()
    if not ok:
        raise RuntimeError(why)
    manifest = json.load(open(os.path.join(REF_ROOT, "MANIFEST.json")))["files"]
    for rel, digest in manifest.items():
        with open(os.path.join(REF_ROOT, rel), "rb") as fh:
            if hashlib.sha256(fh.read()).hexdigest() != digest:
                raise RuntimeError(f"baseline/_ref/{rel} differs from the staged manifest: the reference arm must run unmodified code")
    if "torch_nerf" in sys.modules and not getattr(sys.modules["torch_nerf"], "__file__", "").startswith(REF_ROOT):
        raise RuntimeError("another `torch_nerf` package is already imported")
    sys.path.insert(0, REF_ROOT)
    try:
        m = {
            "NeRF": importlib.import_module("torch_nerf.src.network.nerf").NeRF,
            "PerspectiveCamera": importlib.import_module("torch_nerf.src.renderer.cameras").PerspectiveCamera,
            "QuadratureIntegrator": importlib.import_module(
                "torch_nerf.src.renderer.integrators.quadrature_integrator").QuadratureIntegrator,
            "StratifiedSampler": importlib.import_module(
                "torch_nerf.src.renderer.ray_samplers.stratified_sampler").StratifiedSampler,
            "VolumeRenderer": importlib.import_module("torch_nerf.src.renderer.volume_renderer").VolumeRenderer,
            "PrimitiveCube": importlib.import_module("torch_nerf.src.scene.primitives.cube").PrimitiveCube,
            "PositionalEncoder": importlib.import_module("torch_nerf.src.signal_encoder.positional_encoder").PositionalEncoder,
        }
    finally:
        sys.path.remove(REF_ROOT)
    _mods = m
    return m


class RefSession:
    """The objects one reference training / rendering session holds (runner_utils.py:526-733), on `device`."""

    def __init__(self, device, num_coarse: int = 64, num_fine: int = 128, init_lr: float = 5e-4, end_lr: float = 5e-5,
                 num_iter: int = 300000, seed: int = 0):
        m = load()
        self.m = m
        self.device = torch.device(device)
        # the reference passes an int CUDA index around (train.py:178); "cpu" works for the same argument on the host
        self.dev_arg = torch.cuda.current_device() if self.device.type == "cuda" else "cpu"
        self.sc, self.sf = num_coarse, num_fine
        torch.manual_seed(seed)
        # runner_utils.py:526-550
        import contextlib
        import io

        with contextlib.redirect_stdout(io.StringIO()):  # the constructor prints a warning when camera is None
            self.renderer = m["VolumeRenderer"](m["QuadratureIntegrator"](), m["StratifiedSampler"]())
        # runner_utils.py:569-660: ONE encoder dict shared by both scenes, NeRF(63, 27) x 2
        coord_enc, dir_enc = m["PositionalEncoder"](3, 10, True), m["PositionalEncoder"](3, 4, True)
        enc = {"coord_enc": coord_enc, "dir_enc": dir_enc}
        self.default_net = m["NeRF"](coord_enc.out_dim, dir_enc.out_dim).to(self.device)
        self.fine_net = m["NeRF"](coord_enc.out_dim, dir_enc.out_dim).to(self.device)
        self.default_scene = m["PrimitiveCube"](self.default_net, enc)
        self.fine_scene = m["PrimitiveCube"](self.fine_net, enc)
        # runner_utils.py:663-715
        params = list(self.default_net.parameters()) + list(self.fine_net.parameters())
        self.optimizer = torch.optim.Adam(params, lr=init_lr, eps=1e-8)
        self.scheduler = torch.optim.lr_scheduler.ExponentialLR(self.optimizer, pow(end_lr / init_lr, 1 / num_iter))
        self.loss_func = torch.nn.MSELoss()  # runner_utils.py:731

    def set_camera(self, intrinsic, extrinsic, t_near: float, t_far: float):
        self.renderer.camera = self.m["PerspectiveCamera"](intrinsic, extrinsic, t_near, t_far)

    def train_step(self, pixel_gt: torch.Tensor, intrinsic, extrinsic, t_near: float, t_far: float, num_pixels: int,
                   project_to_ndc: bool = False, pixel_indices: Optional[torch.Tensor] = None):
        """train.py:130-218 for one (image, pose) batch.  pixel_gt (H*W, 3) on the host like the DataLoader's."""
        self.optimizer.zero_grad()
        self.set_camera(intrinsic, extrinsic, t_near, t_far)
        coarse_pred, coarse_indices, coarse_weights = self.renderer.render_scene(
            self.default_scene, num_pixels=num_pixels, num_samples=self.sc, project_to_ndc=project_to_ndc,
            pixel_indices=pixel_indices, device=self.dev_arg)
        coarse_loss = self.loss_func(pixel_gt[coarse_indices, ...].to(self.device), coarse_pred)
        fine_pred, fine_indices, _ = self.renderer.render_scene(
            self.fine_scene, num_pixels=num_pixels, num_samples=(self.sc, self.sf), project_to_ndc=project_to_ndc,
            pixel_indices=coarse_indices, weights=coarse_weights, device=self.dev_arg)
        fine_loss = self.loss_func(pixel_gt[fine_indices, ...].to(self.device), fine_pred)
        loss = coarse_loss + fine_loss
        loss_values = (coarse_loss.item(), fine_loss.item(), loss.item())  # the three .item() syncs of train.py:183-210
        loss.backward()
        self.optimizer.step()
        self.scheduler.step()
        return loss_values

    @torch.no_grad()
    def render_frame(self, intrinsic, extrinsic, t_near: float, t_far: float, img_res: Tuple[int, int], num_pixels: int = 4096,
                     project_to_ndc: bool = False) -> torch.Tensor:
        """render.py:58-107: whole frame, coarse then fine, num_ray_batch = H*W // num_pixels; returns (3, H, W)."""
        self.set_camera(intrinsic, extrinsic, t_near, t_far)
        h, w = img_res
        total = h * w
        img, idx, wts = self.renderer.render_scene(
            self.default_scene, num_pixels=total, num_samples=self.sc, project_to_ndc=project_to_ndc, device=self.dev_arg,
            num_ray_batch=total // num_pixels)
        img, _, _ = self.renderer.render_scene(
            self.fine_scene, num_pixels=total, num_samples=(self.sc, self.sf), project_to_ndc=project_to_ndc,
            pixel_indices=idx, weights=wts, device=self.dev_arg, num_ray_batch=total // num_pixels)
        img = img.reshape(h, w, -1).permute(2, 0, 1)
        return torch.clamp(img, 0.0, 1.0)
