"""GPU test of the drop-in boundary through the REFERENCE's own caller (VERDICT r1 "what's missing" #3): the unmodified
reference `VolumeRenderer` (baseline/_ref, staged by baseline/stage_reference.py) drives a coarse + fine pass and a
backward with this package's sampler / integrator / scene (network + encoders) swapped in ONE AT A TIME and all together,
exactly the way `runners/runner_utils.py:526-550, 569-660` would construct them; every variant must reproduce the
all-reference result on torch-CUDA on identical parameters and identical uniform draws.

Tolerance: 1e-3 max-abs on pixel colours and weights (north-star fp32 gate); gradients 5e-3 of the tensor's max.
Skipped when baseline/_ref has not been staged (it is git-ignored; `__graft_entry__.build()` stages it in the dev container).
"""
from contextlib import contextmanager

import numpy as np
import pytest
import torch

from oracle import nerf_oracle as orc

pytestmark = pytest.mark.gpu
SC, SF = 64, 128


@pytest.fixture(scope="module")
def env():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from baseline import ref_harness as rh

    ok, why = rh.available()
    if not ok:
        pytest.skip(why)
    import torch_nerf_b200 as tn

    tn._lib.load()
    return tn, rh.load()


@contextmanager
def replay_uniforms(queue):
    """Feeds the given CUDA tensors, in order, to torch.rand / torch.rand_like (reference and b200 samplers alike)."""
    q = list(queue)
    orig_rand, orig_rand_like = torch.rand, torch.rand_like

    def fake_rand(*size, **kw):
        t = q.pop(0)
        shape = tuple(size[0]) if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else tuple(size)
        assert tuple(t.shape) == shape, (t.shape, shape)
        return t.clone()

    def fake_rand_like(x, **kw):
        t = q.pop(0)
        assert t.shape == x.shape, (t.shape, x.shape)
        return t.clone()

    torch.rand, torch.rand_like = fake_rand, fake_rand_like
    try:
        yield
    finally:
        torch.rand, torch.rand_like = orig_rand, orig_rand_like
    assert not q, "unused uniforms"


def build(tn, ref, swap):
    """(renderer, coarse scene, fine scene) with the components named in `swap` taken from torch_nerf_b200."""
    integ = tn.QuadratureIntegrator() if "integrator" in swap else ref["QuadratureIntegrator"]()
    sampler = tn.StratifiedSampler() if "sampler" in swap else ref["StratifiedSampler"]()
    import contextlib, io

    with contextlib.redirect_stdout(io.StringIO()):
        renderer = ref["VolumeRenderer"](integ, sampler)
    scenes = []
    for seed in (81, 82):
        params = {k: torch.from_numpy(v.copy()) for k, v in orc.init_nerf_params(seed=seed).items()}
        if "scene" in swap:
            net = tn.NeRF(63, 27, precision="fp32")
            enc = {"coord_enc": tn.PositionalEncoder(3, 10, True), "dir_enc": tn.PositionalEncoder(3, 4, True)}
            cube = tn.PrimitiveCube
        else:
            net = ref["NeRF"](63, 27)
            enc = {"coord_enc": ref["PositionalEncoder"](3, 10, True), "dir_enc": ref["PositionalEncoder"](3, 4, True)}
            cube = ref["PrimitiveCube"]
        net.load_state_dict(params)
        scenes.append(cube(net.cuda(), enc))
    return renderer, scenes[0], scenes[1]


def run(tn, ref, swap, u, pix, target):
    renderer, coarse, fine = build(tn, ref, swap)
    h = w = 64
    focal = orc.blender_focal(w)
    renderer.camera = ref["PerspectiveCamera"]({"f_x": focal, "f_y": focal, "img_width": w, "img_height": h},
                                               torch.from_numpy(orc.pose_spherical(30.0, -30.0, 4.0)), 2.0, 6.0)
    dev = torch.cuda.current_device()
    n = pix.shape[0]
    with replay_uniforms(u):
        rgb_c, idx, w_c = renderer.render_scene(coarse, num_pixels=n, num_samples=SC, project_to_ndc=False, pixel_indices=pix,
                                                device=dev)
        w_c_out = w_c.detach().clone()
        rgb_f, _, w_f = renderer.render_scene(fine, num_pixels=n, num_samples=(SC, SF), project_to_ndc=False, pixel_indices=idx,
                                              weights=w_c, device=dev)
    loss = torch.nn.functional.mse_loss(rgb_c, target) + torch.nn.functional.mse_loss(rgb_f, target)
    loss.backward()
    torch.cuda.synchronize()
    grads = {f"{tag}/{k}": p.grad.detach().cpu().numpy() for tag, sc in (("c", coarse), ("f", fine))
             for k, p in sc.radiance_field.named_parameters()}
    return {"rgb_c": rgb_c.detach().cpu().numpy(), "rgb_f": rgb_f.detach().cpu().numpy(), "w_c": w_c_out.cpu().numpy(),
            "w_f": w_f.detach().cpu().numpy(), "loss": float(loss), "grads": grads}


@pytest.mark.parametrize("swap", [("sampler",), ("integrator",), ("scene",), ("sampler", "integrator", "scene")],
                         ids=["sampler", "integrator", "scene", "all"])
def test_reference_volume_renderer_with_b200_components(env, swap):
    tn, ref = env
    n = 384
    gen = torch.Generator(device="cuda").manual_seed(11)
    u = [torch.rand((n, k), device="cuda", generator=gen) for k in (SC, SC, SF, SF)]
    pix = torch.randperm(64 * 64, generator=torch.Generator().manual_seed(3))[:n]
    target = torch.rand((n, 3), device="cuda", generator=gen)
    base = run(tn, ref, (), u, pix, target)
    got = run(tn, ref, swap, u, pix, target)
    for k in ("rgb_c", "w_c", "rgb_f", "w_f"):
        assert np.abs(got[k] - base[k]).max() <= 1e-3, (k, np.abs(got[k] - base[k]).max())
    assert abs(got["loss"] - base["loss"]) <= 1e-4 * max(1.0, abs(base["loss"]))
    for k, g in base["grads"].items():
        scale = np.abs(g).max() + 1e-12
        assert np.abs(got["grads"][k] - g).max() <= 5e-3 * scale, (k, np.abs(got["grads"][k] - g).max() / scale)
