"""
Builds two tiny synthetic datasets in the reference's on-disk formats under tests/golden/data/ and freezes what the
UNMODIFIED reference loaders (imported from /root/reference) return for them.  Run in the dev container only:

    python tests/golden/make_golden_data.py

The reference reads images with `imageio`, which is not installed here; a three-line stand-in module that decodes with
PIL is registered under that name before the reference modules are imported (the reference code itself is untouched).
"""
import json
import os
import sys
import types

import numpy as np
import torch
from PIL import Image

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(HERE, "data")

stub = types.ModuleType("imageio")
stub.imread = lambda path, **kw: np.asarray(Image.open(path))
sys.modules["imageio"] = stub
sys.path.insert(0, REF)

from torch_nerf.src.utils.data.blender_dataset import NeRFBlenderDataset  # noqa: E402
from torch_nerf.src.utils.data.llff_dataset import LLFFDataset  # noqa: E402


def random_pose(rng):
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    m = np.eye(4)
    m[:3, :3] = q
    m[:3, 3] = rng.normal(size=3) * 2.0
    return m


def make_blender(rng):
    root = os.path.join(DATA, "blender", "lego")
    for split, count in (("train", 3), ("val", 2), ("test", 4)):
        os.makedirs(os.path.join(root, split), exist_ok=True)
        frames = []
        for i in range(count):
            img = rng.integers(0, 256, size=(12, 16, 4), dtype=np.uint8)
            img[:4, :5, 3] = 0          # fully transparent corner -> white under white_bg
            img[8:, 10:, 3] = 255
            Image.fromarray(img, "RGBA").save(os.path.join(root, split, f"r_{i}.png"))
            frames.append({"file_path": f"./{split}/r_{i}", "rotation": 0.012, "transform_matrix": random_pose(rng).tolist()})
        with open(os.path.join(root, f"transforms_{split}.json"), "w") as fh:
            json.dump({"camera_angle_x": 0.6911112070083618, "frames": frames}, fh, indent=1)


def make_llff(rng):
    root = os.path.join(DATA, "llff", "fern")
    n = 6
    os.makedirs(os.path.join(root, "images"), exist_ok=True)
    os.makedirs(os.path.join(root, "images_2"), exist_ok=True)
    rows = []
    for i in range(n):
        Image.fromarray(rng.integers(0, 256, size=(12, 16, 3), dtype=np.uint8), "RGB").save(os.path.join(root, "images", f"img_{i:02d}.png"))
        Image.fromarray(rng.integers(0, 256, size=(6, 8, 3), dtype=np.uint8), "RGB").save(os.path.join(root, "images_2", f"img_{i:02d}.png"))
        pose = random_pose(rng)[:3, :4]
        pose[:3, :3] = np.eye(3) + 0.1 * rng.normal(size=(3, 3))   # forward-facing-ish cameras
        hwf = np.array([[12.0], [16.0], [20.0]])
        rows.append(np.concatenate([np.concatenate([pose, hwf], 1).reshape(-1), [1.5 + rng.random(), 9.0 + rng.random()]]))
    np.save(os.path.join(root, "poses_bounds.npy"), np.stack(rows, 0))


def main():
    rng = np.random.default_rng(2024)
    make_blender(rng)
    make_llff(rng)
    out = {}
    for split, half in (("train", False), ("test", True)):
        ds = NeRFBlenderDataset(os.path.join(DATA, "blender"), "lego", split, half_res=half, white_bg=True)
        tag = f"blender_{split}{'_half' if half else ''}"
        items = [ds[i] for i in range(len(ds))]
        out[tag + "/imgs"] = torch.stack([x[0] for x in items]).numpy()
        out[tag + "/poses"] = torch.stack([x[1] for x in items]).numpy()
        out[tag + "/cam"] = np.array([ds.img_height, ds.img_width, ds.focal_length], dtype=np.float64)
        out[tag + "/render_poses"] = ds.render_poses.numpy()
    ds = NeRFBlenderDataset(os.path.join(DATA, "blender"), "lego", "val", half_res=False, white_bg=False)
    out["blender_val_nobg/imgs"] = torch.stack([ds[i][0] for i in range(len(ds))]).numpy()
    for tag, kw in (("llff", dict(recenter=True, bd_factor=0.75, spherify=False)),
                    ("llff_spherify", dict(recenter=True, bd_factor=0.75, spherify=True)),
                    ("llff_raw", dict(recenter=False, bd_factor=None, spherify=False))):
        ds = LLFFDataset(os.path.join(DATA, "llff"), "fern", factor=2, **kw)
        out[tag + "/imgs"] = ds._imgs.numpy()
        out[tag + "/poses"] = ds._poses.numpy()
        out[tag + "/cam"] = np.array([ds.img_height, ds.img_width, ds.focal_length], dtype=np.float64)
        out[tag + "/z_bounds"] = ds.z_bounds.numpy()
        out[tag + "/render_poses"] = ds.render_poses.numpy()
        out[tag + "/idx_test"] = np.array(ds._idx_test)
    # (path_zflat=True cannot be frozen: load_llff.py:535 makes the key-frame count a float, which numpy >= 1.18 rejects)
    np.savez_compressed(os.path.join(HERE, "datasets.npz"), **out)
    print("wrote", os.path.join(HERE, "datasets.npz"), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
