"""CPU tests of the dataset readers against what the reference's loaders returned for the
tiny on-disk fixtures under tests/golden/data (frozen by tests/golden/make_golden_data.py)."""
import os

import numpy as np
import pytest
import torch

from conftest import load_golden

from torch_nerf_b200 import datasets as ds

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "data")


@pytest.fixture(scope="module")
def g():
    return load_golden("datasets.npz")


@pytest.mark.parametrize("split,half,tag", [("train", False, "blender_train"), ("test", True, "blender_test_half")])
def test_blender_dataset_matches_reference(g, split, half, tag):
    d = ds.BlenderDataset(os.path.join(DATA, "blender"), "lego", split, half_res=half, white_bg=True)
    assert len(d) == g[tag + "/imgs"].shape[0]
    imgs = torch.stack([d[i][0] for i in range(len(d))]).numpy()
    poses = torch.stack([d[i][1] for i in range(len(d))]).numpy()
    np.testing.assert_array_equal(imgs, g[tag + "/imgs"])          # same decoder output, same resize kernel
    np.testing.assert_array_equal(poses, g[tag + "/poses"])
    np.testing.assert_allclose([d.img_height, d.img_width, d.focal_length], g[tag + "/cam"], rtol=1e-12)
    np.testing.assert_allclose(d.render_poses.numpy(), g[tag + "/render_poses"], rtol=0, atol=1e-6)
    assert imgs.dtype == np.float32 and imgs.shape[-1] == 3


def test_blender_white_background_is_exact_alpha_zero_only(g):
    d = ds.BlenderDataset(os.path.join(DATA, "blender"), "lego", "val", half_res=False, white_bg=False)
    imgs = torch.stack([d[i][0] for i in range(len(d))]).numpy()
    np.testing.assert_array_equal(imgs, g["blender_val_nobg/imgs"])
    w = ds.BlenderDataset(os.path.join(DATA, "blender"), "lego", "val", half_res=False, white_bg=True)
    assert float(w[0][0][:4, :5].min()) == 1.0                      # the transparent corner turned white
    assert not np.array_equal(w[0][0].numpy(), imgs[0])


def test_blender_argument_errors():
    with pytest.raises(ValueError):
        ds.BlenderDataset(os.path.join(DATA, "blender"), "lego", "training", half_res=False)
    with pytest.raises(ValueError):
        ds.BlenderDataset(os.path.join(DATA, "blender"), "teapot", "train", half_res=False)
    with pytest.raises(ValueError):
        ds.load_blender_data(os.path.join(DATA, "blender", "lego"), "dev")


@pytest.mark.parametrize("tag,kw", [("llff", dict(recenter=True, bd_factor=0.75, spherify=False)),
                                    ("llff_spherify", dict(recenter=True, bd_factor=0.75, spherify=True)),
                                    ("llff_raw", dict(recenter=False, bd_factor=None, spherify=False))])
def test_llff_dataset_matches_reference(g, tag, kw):
    d = ds.LLFFDataset(os.path.join(DATA, "llff"), "fern", factor=2, **kw)
    np.testing.assert_array_equal(d._imgs.numpy(), g[tag + "/imgs"])
    np.testing.assert_allclose(d._poses.numpy(), g[tag + "/poses"], rtol=0, atol=2e-6)
    np.testing.assert_allclose([d.img_height, d.img_width, d.focal_length], g[tag + "/cam"], rtol=1e-7)
    np.testing.assert_allclose(d.z_bounds.numpy(), g[tag + "/z_bounds"], rtol=1e-6)
    np.testing.assert_allclose(d.render_poses.numpy(), g[tag + "/render_poses"], rtol=0, atol=5e-6)
    assert d._idx_test == int(g[tag + "/idx_test"])
    img, pose = d[1]
    assert tuple(img.shape) == (6, 8, 3) and tuple(pose.shape) == tuple(g[tag + "/poses"].shape[1:])


def test_llff_errors_and_zflat_path():
    with pytest.raises(ValueError):
        ds.LLFFDataset(os.path.join(DATA, "llff"), "garden", 2, True, 0.75, False)
    with pytest.raises(ValueError):  # no images_4 directory: the reference would shell out to mogrify, we refuse
        ds.load_llff_data(os.path.join(DATA, "llff", "fern"), factor=4)
    r = ds.load_llff_data(os.path.join(DATA, "llff", "fern"), factor=2, path_zflat=True)
    assert r[4].shape == (60, 3, 4) and np.isfinite(r[4]).all()   # (the reference's zflat path no longer runs on numpy >= 1.18)
