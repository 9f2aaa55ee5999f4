"""On-disk dataset formats the reference trains from, host-side numpy only.

    Blender ("nerf_synthetic"):  transforms_{split}.json + RGBA PNGs      ref src/utils/data/load_blender.py:112-190,
                                                                           blender_dataset.py:22-102
    LLFF ("nerf_llff_data"):     poses_bounds.npy + images_{factor}/*.png  ref src/utils/data/load_llff.py:88-193,461-570,
                                                                           llff_dataset.py:22-100

Both classes yield (pixel_gt (H, W, 3) float32 in [0,1], camera-to-world pose) per view, which is what
Trainer.train_one_epoch consumes.  PNGs are read with PIL (the reference uses imageio, which is not a dependency here);
the half-resolution Blender path uses the same cv2.INTER_AREA resize as the reference.  Unlike the reference, a missing
``images_{factor}`` directory is an error: the reference shells out to ImageMagick's mogrify to create it."""
from __future__ import annotations

import json
import os
from pathlib import Path
from typing import List, Sequence, Tuple

import numpy as np
import torch

_IMG_EXT = ("JPG", "jpg", "png")


def _read_image(path) -> np.ndarray:
    from PIL import Image

    with Image.open(path) as im:
        return np.asarray(im)


def pose_spherical(theta_deg: float, phi_deg: float, radius: float) -> np.ndarray:
    """load_blender.py:78-109: translate along z by `radius`, rotate by phi about x, by theta about y, then swap the
    axes into Blender's convention.  (4,4) float32."""
    f32 = np.float32
    phi, th = phi_deg / 180.0 * np.pi, theta_deg / 180.0 * np.pi
    trans = np.eye(4, dtype=f32)
    trans[2, 3] = radius
    rot_phi = np.array([[1, 0, 0, 0], [0, np.cos(phi), -np.sin(phi), 0], [0, np.sin(phi), np.cos(phi), 0], [0, 0, 0, 1]], dtype=f32)
    rot_th = np.array([[np.cos(th), 0, -np.sin(th), 0], [0, 1, 0, 0], [np.sin(th), 0, np.cos(th), 0], [0, 0, 0, 1]], dtype=f32)
    swap = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=f32)
    return (swap @ (rot_th @ (rot_phi @ trans))).astype(f32)


# ------------------------------------------------------------------------------------------------ Blender
def load_blender_data(base_dir, dataset_type: str, half_res: bool = False, test_idx_skip: int = 1):
    """load_blender.py:112-190.  Returns (imgs (B,H,W,4) f32, poses (B,4,4) f32, [H, W, focal], render_poses (40,4,4),
    image file stems)."""
    if dataset_type not in ("train", "val", "test"):
        raise ValueError(f"Unsupported dataset type. Expected one of ['train', 'val', 'test']. Got {dataset_type}")
    base_dir = Path(base_dir)
    with open(base_dir / f"transforms_{dataset_type}.json", "r") as fh:
        meta = json.load(fh)
    skip = 1 if (dataset_type == "train" or test_idx_skip == 0) else test_idx_skip
    frames = meta["frames"][::skip]
    files = [base_dir / f"{fr['file_path']}.png" for fr in frames]
    imgs = (np.array([_read_image(f) for f in files]) / 255.0).astype(np.float32)
    poses = np.array([fr["transform_matrix"] for fr in frames]).astype(np.float32)
    height, width = imgs[0].shape[:2]
    focal = float(0.5 * width / np.tan(0.5 * float(meta["camera_angle_x"])))
    render_poses = np.stack([pose_spherical(a, -30.0, 4.0) for a in np.linspace(-180, 180, 41)[:-1]], 0)
    if half_res:
        import cv2

        height, width, focal = height // 2, width // 2, focal / 2.0
        imgs = np.stack([cv2.resize(im, (width, height), interpolation=cv2.INTER_AREA) for im in imgs], 0).astype(np.float32)
    return imgs, poses, [height, width, focal], render_poses, [f.stem for f in files]


class BlenderDataset(torch.utils.data.Dataset):
    """blender_dataset.py:22-122 (same constructor arguments, properties and items)."""

    SCENES = ("chair", "drums", "ficus", "hotdog", "lego", "materials", "mic", "ship")

    def __init__(self, root_dir, scene_name: str, data_type: str, half_res: bool, white_bg: bool = True):
        if data_type not in ("train", "val", "test"):
            raise ValueError(f"Unsupported dataset type. Expected one of ['train', 'val', 'test']. Got {data_type}")
        if scene_name not in self.SCENES:
            raise ValueError(f"Unsupported scene type. Expected one of {list(self.SCENES)}. Got {scene_name}.")
        root = Path(root_dir)
        if not root.exists():
            raise ValueError(f"The directory {root} does not exist.")
        super().__init__()
        self._white_bg = white_bg
        self._imgs, self._poses, cam, render_poses, self._img_fnames = load_blender_data(root / scene_name, data_type, half_res)
        self._img_height, self._img_width, self._focal_length = cam
        self._render_poses = torch.from_numpy(render_poses)
        if self._imgs.shape[0] != self._poses.shape[0]:
            raise AssertionError(f"Dataset sizes do not match. Got {self._imgs.shape[0]} images and {self._poses.shape[0]} camera poses.")

    def __len__(self) -> int:
        return self._imgs.shape[0]

    def __getitem__(self, index: int) -> Tuple[torch.Tensor, torch.Tensor]:
        img = torch.tensor(self._imgs[index])
        if self._white_bg:
            # blender_dataset.py:97-99: pixels whose alpha is EXACTLY zero become white; nothing is alpha-blended
            img[img[..., -1] == 0.0, :] = 1.0
        return img[..., :-1], torch.tensor(self._poses[index])

    img_height = property(lambda self: self._img_height)
    img_width = property(lambda self: self._img_width)
    focal_length = property(lambda self: self._focal_length)
    render_poses = property(lambda self: self._render_poses)


# ------------------------------------------------------------------------------------------------ LLFF
def _unit(v: np.ndarray) -> np.ndarray:
    return v / np.linalg.norm(v)


def _look_along(z_axis, up, position) -> np.ndarray:
    """load_llff.py:231-260: (3,4) pose whose third column is `z_axis`, with `up` fixing the roll."""
    z = _unit(z_axis)
    x = _unit(np.cross(up, z))
    y = _unit(np.cross(z, x))
    return np.stack([x, y, z, position], 1)


def average_pose(poses: np.ndarray) -> np.ndarray:
    """load_llff.py:284-312: mean position, summed viewing axis, summed y axis as the up hint."""
    return _look_along(_unit(poses[:, :3, 2].sum(0)), poses[:, :3, 1].sum(0), poses[:, :3, 3].mean(0))


def recenter_poses(poses: np.ndarray) -> np.ndarray:
    """load_llff.py:353-376: express every pose in the frame of the average pose."""
    last_row = np.array([[0.0, 0.0, 0.0, 1.0]])
    center = np.concatenate([average_pose(poses), last_row], 0)
    homog = np.concatenate([poses[:, :3, :4], np.broadcast_to(last_row, (poses.shape[0], 1, 4))], 1)
    out = poses.copy()
    out[:, :3, :4] = (np.linalg.inv(center) @ homog)[:, :3, :4]
    return out


def spiral_path(center_pose: np.ndarray, up: np.ndarray, radii: Sequence[float], focus: float, z_rate: float, rotations: int,
                count) -> List[np.ndarray]:
    """load_llff.py:315-350: `count` poses on a spiral around `center_pose`, all looking at the focus point."""
    scale = np.array(list(radii) + [1.0])
    look_at = center_pose[:3, :4] @ np.array([0.0, 0.0, -focus, 1.0])
    out = []
    for theta in np.linspace(0.0, 2.0 * np.pi * rotations, int(count) + 1)[:-1]:
        pos = center_pose[:3, :4] @ (np.array([np.cos(theta), -np.sin(theta), -np.sin(theta * z_rate), 1.0]) * scale)
        out.append(_look_along(pos - look_at, up, pos))
    return out


def spherify_poses(poses: np.ndarray, bounds: np.ndarray):
    """load_llff.py:382-458: recentre on the point closest to all optical axes, rescale to the unit sphere, and build a
    120-pose circular render path.  Returns (poses (N,3,5), render_poses (120,3,5), bounds)."""
    def to44(p):
        return np.concatenate([p, np.broadcast_to(np.array([[[0.0, 0.0, 0.0, 1.0]]]), (p.shape[0], 1, 4))], 1)

    d, o = poses[:, :3, 2:3], poses[:, :3, 3:4]
    proj = np.eye(3) - d * np.transpose(d, (0, 2, 1))
    focus_pt = np.squeeze(-np.linalg.inv((np.transpose(proj, (0, 2, 1)) @ proj).mean(0)) @ (-proj @ o).mean(0))
    up = _unit((poses[:, :3, 3] - focus_pt).mean(0))
    ax1 = _unit(np.cross([0.1, 0.2, 0.3], up))
    ax2 = _unit(np.cross(up, ax1))
    frame = np.stack([ax1, ax2, up, focus_pt], 1)
    reset = np.linalg.inv(to44(frame[None])) @ to44(poses[:, :3, :4])
    radius = np.sqrt(np.mean(np.sum(np.square(reset[:, :3, 3]), -1)))
    s = 1.0 / radius
    reset[:, :3, 3] *= s
    bounds *= s
    radius *= s
    height = np.mean(reset[:, :3, 3], 0)[2]
    circle = np.sqrt(radius ** 2 - height ** 2)
    ring = []
    for th in np.linspace(0.0, 2.0 * np.pi, 120):
        origin = np.array([circle * np.cos(th), circle * np.sin(th), height])
        z = _unit(origin)
        x = _unit(np.cross(z, np.array([0.0, 0.0, -1.0])))
        y = _unit(np.cross(z, x))
        ring.append(np.stack([x, y, z, origin], 1))
    ring = np.stack(ring, 0)
    hwf = poses[0, :3, -1:]
    ring = np.concatenate([ring, np.broadcast_to(hwf, ring[:, :3, -1:].shape)], -1)
    reset = np.concatenate([reset[:, :3, :4], np.broadcast_to(hwf, reset[:, :3, -1:].shape)], -1)
    return reset, ring, bounds


def _load_llff_raw(base_dir: str, factor):
    """load_llff.py:88-193 (factor path): poses_bounds.npy (N,17) = 3x5 [R | t | (H,W,f)] + (near, far) per view."""
    raw = np.load(os.path.join(base_dir, "poses_bounds.npy"))
    cam = raw[:, :-2].reshape(-1, 3, 5).transpose(1, 2, 0)  # (3, 5, N)
    bounds = raw[:, -2:].transpose(1, 0)
    extr, intr = cam[:, :-1, :], cam[:, -1, :]
    suffix = "" if factor is None else f"_{factor}"
    img_dir = os.path.join(base_dir, "images" + suffix)
    if not os.path.exists(img_dir):
        raise ValueError(f"The base directory of dataset {img_dir} does not exist.")
    files = [os.path.join(img_dir, f) for f in sorted(os.listdir(img_dir)) if f.endswith(_IMG_EXT)]
    if cam.shape[-1] != len(files):
        raise ValueError(f"Mismatch between imgs {len(files)} and poses {cam.shape[-1]}.")
    first = _read_image(files[0])
    intr[:2, :] = np.array(first.shape[:2]).reshape(2, 1)
    intr[2, :] *= 1.0 / (1 if factor is None else factor)
    # LLFF stores [down, right, back]; the renderer wants [right, up, back] (bmild/nerf issue 34)
    extr = np.concatenate([extr[:, 1:2, :], -extr[:, 0:1, :], extr[:, 2:, :]], 1)
    imgs = np.stack([_read_image(f)[..., :3] / 255.0 for f in files], 0).astype(np.float32)
    mv = lambda x: np.moveaxis(x, -1, 0).astype(np.float32)  # noqa: E731
    return imgs, mv(extr), mv(intr), mv(bounds)


def load_llff_data(base_dir: str, factor: int = 8, recenter: bool = True, bd_factor: float = 0.75, spherify: bool = False,
                   path_zflat: bool = False):
    """load_llff.py:461-570.  Returns (imgs (N,H,W,3), extrinsics (N,3,4), intrinsics (N,3) = (H, W, focal),
    z_bounds (N,2), render_poses, index of the hold-out view)."""
    imgs, extr, intr, bounds = _load_llff_raw(base_dir, factor)
    scale = 1.0 if bd_factor is None else 1.0 / (bounds.min() * bd_factor)
    extr[:, :3, 3] *= scale
    bounds *= scale
    if recenter:
        extr = recenter_poses(extr)
    if spherify:
        extr, render_poses, bounds = spherify_poses(extr, bounds)
    else:
        path_center = average_pose(extr)
        up = _unit(extr[:, :, 1].sum(0))
        close, far = bounds.min() * 0.9, bounds.max() * 5.0
        focus = 1.0 / ((1.0 - 0.75) / close + 0.75 / far)
        radii = np.percentile(np.abs(extr[:, :, 3]), 90, 0)
        count, rotations = 120, 2
        if path_zflat:
            path_center[:3, 3] = path_center[:3, 3] + (-close * 0.1) * path_center[:3, 2]
            radii[2] = 0.0
            rotations, count = 1, count / 2
        render_poses = spiral_path(path_center, up, radii, focus, 0.5, rotations, count)
    render_poses = np.array(render_poses).astype(np.float32)
    center = average_pose(extr)
    holdout = int(np.argmin(np.sum(np.square(center[:3, 3] - extr[:, :3, 3]), -1)))
    return imgs.astype(np.float32), extr.astype(np.float32), intr, bounds, render_poses, holdout


class LLFFDataset(torch.utils.data.Dataset):
    """llff_dataset.py:22-134 (same constructor arguments, properties and items)."""

    SCENES = ("fern", "flower", "fortress", "horns", "leaves", "orchids", "room", "trex")

    def __init__(self, root_dir: str, scene_name: str, factor: int, recenter: bool, bd_factor: float, spherify: bool):
        if scene_name not in self.SCENES:
            raise ValueError(f"Unsupported scene type. Expected one of {list(self.SCENES)}. Got {scene_name}.")
        if not os.path.exists(root_dir):
            raise ValueError(f"The directory {root_dir} does not exist.")
        super().__init__()
        imgs, poses, cam, bounds, render_poses, self._idx_test = load_llff_data(
            str(os.path.join(root_dir, scene_name)), factor=factor, recenter=recenter, bd_factor=bd_factor, spherify=spherify)
        self._imgs, self._poses = torch.tensor(imgs), torch.tensor(poses)
        self._z_bounds, self._render_poses = torch.tensor(bounds), torch.tensor(render_poses)
        self._img_height, self._img_width, self._focal_length = int(cam[0, 0]), int(cam[0, 1]), float(cam[0, 2])
        if self._imgs.shape[0] != self._poses.shape[0]:
            raise AssertionError(f"Dataset sizes do not match. Got {self._imgs.shape[0]} images and {self._poses.shape[0]} camera poses.")

    def __len__(self) -> int:
        return self._imgs.shape[0]

    def __getitem__(self, index: int) -> Tuple[torch.Tensor, torch.Tensor]:
        return self._imgs[index], self._poses[index]

    img_height = property(lambda self: self._img_height)
    img_width = property(lambda self: self._img_width)
    focal_length = property(lambda self: self._focal_length)
    render_poses = property(lambda self: self._render_poses)
    z_bounds = property(lambda self: self._z_bounds)
