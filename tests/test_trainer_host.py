"""CPU tests of the callers around the hot path: pixel selection of the warm-up epochs, the
learning-rate schedule and the reference's checkpoint wire format (flat optimizer state <-> 44-tensor Adam state)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import nerf_oracle as orc

import torch_nerf_b200 as tn
from torch_nerf_b200 import checkpoint as ck
from torch_nerf_b200.engine import FlatParams
from torch_nerf_b200.trainer import center_crop_pixel_indices, exp_lr_gamma, psnr, save_png


@pytest.mark.parametrize("h,w", [(800, 800), (756, 1008), (100, 100), (5, 7)])
def test_center_crop_matches_oracle(h, w):
    got = center_crop_pixel_indices(h, w).numpy()
    np.testing.assert_array_equal(got, orc.center_crop_pixel_indices(h, w))
    if h >= 100:
        rows, cols = got // w, got % w
        assert rows.min() >= 0 and rows.max() < h and cols.min() >= 0 and cols.max() < w
        assert len(np.unique(got)) == len(got)


def test_lr_schedule_matches_reference_formula():
    g = exp_lr_gamma(5e-4, 5e-5, 300000)
    assert g == orc.exp_lr_gamma(5e-4, 5e-5, 300000)
    assert abs(5e-4 * g ** 300000 - 5e-5) < 1e-12


def _nets(seed):
    torch.manual_seed(seed)
    return tn.NeRF(63, 27), tn.NeRF(63, 27)


def _fake_grads(params, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(p.shape, generator=g) * 1e-3 for p in params]


def test_checkpoint_round_trip_with_reference_style_optimizer(tmp_path):
    """A checkpoint written from the flat optimizer must load into the reference's layout -- Adam over the 44
    parameter tensors of (coarse, fine) + ExponentialLR -- and continue to the same parameters, and vice versa."""
    gamma = exp_lr_gamma(5e-4, 5e-5, 1000)
    # --- "ours": flat buffer, one parameter
    c, f = _nets(1)
    flat = FlatParams([c, f])
    opt = torch.optim.Adam([flat.param], lr=5e-4, eps=1e-8)
    sch = torch.optim.lr_scheduler.ExponentialLR(opt, gamma)
    per_tensor = [p for n in (c, f) for p in n.ordered_parameters()]
    for it in range(3):
        for p, g in zip(per_tensor, _fake_grads(per_tensor, 100 + it)):
            p.grad.copy_(g)  # views of the flat gradient buffer
        opt.step(); sch.step()
    path = ck.save_ckpt(tmp_path, 7, c, f, opt, sch, flat)
    assert os.path.basename(path) == "ckpt_000007.pth"
    raw = torch.load(path, map_location="cpu")
    assert set(raw) == {"epoch", "optimizer_state_dict", "scheduler_state_dict", "scene_default", "scene_fine"}
    assert list(raw["scene_default"]) == [f"{n}.{t}" for n in ("fc_in", "fc_1", "fc_2", "fc_3", "fc_4", "fc_5", "fc_6",
                                                               "fc_7", "fc_8", "fc_9", "fc_out") for t in ("weight", "bias")]
    assert len(raw["optimizer_state_dict"]["state"]) == 44 and raw["optimizer_state_dict"]["param_groups"][0]["params"] == list(range(44))
    # --- "reference": runner_utils.py:684-717 + 786-831, with the reference's own NeRF class when it is available
    ref_cls = None
    if os.path.isdir("/root/reference/torch_nerf"):
        sys.path.insert(0, "/root/reference")
        try:
            from torch_nerf.src.network.nerf import NeRF as ref_cls  # noqa: N813
        except Exception:  # pragma: no cover
            ref_cls = None
    rc, rf = (ref_cls(63, 27), ref_cls(63, 27)) if ref_cls is not None else _nets(99)
    ropt = torch.optim.Adam(list(rc.parameters()) + list(rf.parameters()), lr=5e-4, eps=1e-8)
    rsch = torch.optim.lr_scheduler.ExponentialLR(ropt, gamma)
    rc.load_state_dict(raw["scene_default"]); rf.load_state_dict(raw["scene_fine"])
    ropt.load_state_dict(raw["optimizer_state_dict"]); rsch.load_state_dict(raw["scheduler_state_dict"])
    assert raw["epoch"] == 7
    # one more identical step on both sides
    grads = _fake_grads(per_tensor, 555)
    for p, g in zip(per_tensor, grads):
        p.grad.copy_(g)
    opt.step(); sch.step()
    for p, g in zip(list(rc.parameters()) + list(rf.parameters()), grads):
        p.grad = g.clone()
    ropt.step(); rsch.step()
    for a, b in zip(per_tensor, list(rc.parameters()) + list(rf.parameters())):
        np.testing.assert_allclose(a.detach().numpy(), b.detach().numpy(), rtol=0, atol=1e-7)
    assert opt.param_groups[0]["lr"] == pytest.approx(ropt.param_groups[0]["lr"], rel=1e-12)
    # --- and back: a checkpoint in the reference's native form loads into a fresh flat trainer state
    ref_dir = tmp_path / "ref"
    os.makedirs(ref_dir)
    torch.save({"epoch": 8, "optimizer_state_dict": ropt.state_dict(), "scheduler_state_dict": rsch.state_dict(),
                "scene_default": rc.state_dict(), "scene_fine": rf.state_dict()}, ck.ckpt_path(ref_dir, 8))
    c2, f2 = _nets(2)
    flat2 = FlatParams([c2, f2])
    opt2 = torch.optim.Adam([flat2.param], lr=1e-3, eps=1e-8)
    sch2 = torch.optim.lr_scheduler.ExponentialLR(opt2, gamma)
    assert ck.load_ckpt(ref_dir, c2, f2, opt2, sch2, flat2) == 8
    per2 = [p for n in (c2, f2) for p in n.ordered_parameters()]
    assert per2[0].data_ptr() == flat2.flat.data_ptr()  # still aliases the flat buffer
    grads = _fake_grads(per_tensor, 777)
    for p, g in zip(per2, grads):
        p.grad.copy_(g)
    opt2.step(); sch2.step()
    for p, g in zip(list(rc.parameters()) + list(rf.parameters()), grads):
        p.grad = g.clone()
    ropt.step(); rsch.step()
    for a, b in zip(per2, list(rc.parameters()) + list(rf.parameters())):
        np.testing.assert_allclose(a.detach().numpy(), b.detach().numpy(), rtol=0, atol=1e-7)


def test_adam_matches_oracle_restatement():
    """torch's Adam on the flat buffer against the numpy restatement (first two updates)."""
    c, f = _nets(3)
    flat = FlatParams([c, f])
    opt = torch.optim.Adam([flat.param], lr=5e-4, eps=1e-8)
    p0 = flat.flat.clone().numpy()
    m = np.zeros_like(p0); v = np.zeros_like(p0); p = p0
    for step in (1, 2):
        g = torch.randn(flat.flat.shape, generator=torch.Generator().manual_seed(step)) * 1e-3
        flat.grad.copy_(g)
        opt.step()
        p, m, v = orc.adam_step(p, g.numpy(), m, v, step, 5e-4)
        np.testing.assert_allclose(flat.flat.detach().numpy(), p, rtol=0, atol=2e-7)


def test_load_ckpt_missing_dir_returns_zero(tmp_path):
    c, f = _nets(4)
    assert ck.load_ckpt(tmp_path / "nope", c, f) == 0
    os.makedirs(tmp_path / "empty")
    assert ck.load_ckpt(tmp_path / "empty", c, f) == 0


def test_psnr_and_png(tmp_path):
    a = torch.rand(3, 8, 9)
    b = (a + 0.1).clamp(0, 1)
    expect = 10 * np.log10(1.0 / float(((a - b) ** 2).mean()))
    assert float(psnr(a, b)) == pytest.approx(expect, rel=1e-5)
    save_png(a, str(tmp_path / "x.png"))
    from PIL import Image

    back = np.asarray(Image.open(tmp_path / "x.png"))
    assert back.shape == (8, 9, 3)
    np.testing.assert_array_equal(back, (a * 255 + 0.5).clamp(0, 255).permute(1, 2, 0).to(torch.uint8).numpy())
