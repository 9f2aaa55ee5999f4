"""GPU parity at the benchmark's full sizes (VERDICT r1 "what's missing" #4): `HotPathEngine.render_frame` -- the path
bench.py times -- on a whole 800x800 Blender-shaped frame (config C3) and a whole 1008x756 LLFF-shaped NDC frame
(config C5, reference runners/render.py:81-99 with the chunking of volume_renderer.py:229-254), with injected uniform
draws, checked against the oracle on a random 4096-pixel subset (rays are independent, so a subset of rows of the same
uniforms reproduces exactly those pixels).  Row counts reach 1.46e8 and byte offsets 1.8e9 here.

Tolerances: fp32 validation mode <= 1e-3 max-abs on the pixel colour (north-star gate); bf16 tensor-core mode: mean
abs error <= 2e-3 and 99.5th percentile <= 2e-2 when the oracle's fine pass is given the engine's own coarse weights
(same samples, so only the MLP's bf16 rounding differs).  The PSNR gate (bf16 within 0.1 dB of fp32) is scored against a STRUCTURED
target: the fp32 render of a different parameter set at 100x100.
"""
import numpy as np
import pytest
import torch

from oracle import nerf_oracle as orc

pytestmark = pytest.mark.gpu
SC, SF = 64, 128


@pytest.fixture(scope="module")
def tn():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import torch_nerf_b200 as mod

    mod._lib.load()
    return mod


def make_nets(tn, seed_c, seed_f, precision):
    out, params = [], []
    for seed in (seed_c, seed_f):
        p = orc.init_nerf_params(seed=seed)
        net = tn.NeRF(63, 27, precision=precision)
        net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.items()})
        out.append(net.cuda())
        params.append(p)
    return out, params


def frame_config(tn, name):
    if name == "c3_blender_800":
        h = w = 800
        focal = orc.blender_focal(w)
        c2w = orc.pose_spherical(30.0, -30.0, 4.0)
        near, far, ndc = 2.0, 6.0, False
    else:  # config 5: LLFF-shaped forward-facing view, NDC rays, near/far forced to 0/1 (runner_utils.py:489-491)
        w, h, focal = 1008, 756, 815.0
        c2w = np.eye(4, dtype=np.float32)
        c2w[:3, 3] = [0.05, -0.02, 0.1]
        near, far, ndc = 0.0, 1.0, True
    cam = tn.PerspectiveCamera({"f_x": focal, "f_y": focal, "img_width": w, "img_height": h}, torch.from_numpy(c2w), near, far)
    return cam, h, w, focal, c2w, near, far, ndc


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
@pytest.mark.parametrize("config", ["c3_blender_800", "c5_llff_1008x756_ndc"])
def test_render_frame_full_size_vs_oracle_subset(tn, config, precision):
    from torch_nerf_b200.engine import HotPathEngine

    cam, h, w, focal, c2w, near, far, ndc = frame_config(tn, config)
    n = h * w
    (coarse, fine), params = make_nets(tn, 71, 72, precision)
    eng = HotPathEngine(coarse, fine, SC, SF, precision=precision)
    gen = torch.Generator(device="cuda").manual_seed(2024)
    u = tuple(torch.rand((n, k), device="cuda", generator=gen) for k in (SC, SC, SF, SF))
    # fp32 validation mode keeps ~11 KB of activations per sample row: render in chunks like the reference's num_ray_batch
    img = eng.render_frame(cam, ndc, 0, n, max_rays=(1 << 20) if precision == "bf16" else 4096, uniforms=u)
    torch.cuda.synchronize()
    assert img.shape == (n, 3) and bool(torch.isfinite(img).all())
    rng = np.random.default_rng(5)
    sub = np.sort(rng.choice(n, size=4096, replace=False))
    sub[0], sub[-1] = 0, n - 1  # first and last pixel: the extreme row / byte offsets
    sub_t = torch.from_numpy(sub).cuda()
    us = [x[sub_t].cpu().numpy() for x in u]
    coords = orc.screen_coords(h, w)[sub]
    o, d = orc.generate_rays(coords, orc.make_intrinsic(focal, focal, w, h), c2w, near, h, w, ndc)
    co = orc.render_pass(params[0], o, d, near, far, SC, (us[0],))
    if precision == "bf16":
        # single chunk: the engine's coarse pass of the whole frame is still in its workspaces
        w_gpu = eng.last["coarse"]["w"][sub_t].cpu().numpy()
        # (mean, not max: with delta_last = 1e8 the last sample's weight flips between 0 and T_last whenever bf16 rounding
        #  flips the sign of a near-zero density, in the reference arithmetic just as here -- DESIGN.md section 2)
        assert np.abs(w_gpu - co["weights"]).mean() < 1e-3
        w_for_fine = w_gpu.copy()
    else:
        w_for_fine = co["weights"].copy()
    fi = orc.render_pass(params[1], o, d, near, far, (SC, SF), (us[1], us[2], us[3]), weights=w_for_fine)
    got = img[sub_t].cpu().numpy()
    ref = np.clip(fi["rgb"], 0.0, 1.0)
    err = np.abs(got - ref)
    if precision == "fp32":
        assert err.max() <= 1e-3, (err.max(), err.mean())
    else:
        # same caveat for single pixels: bound the mean (VERDICT gate) and the 99.5th percentile, report the max
        q = float(np.quantile(err, 0.995))
        print(f"{config} bf16: mean {err.mean():.2e}  p99.5 {q:.2e}  max {err.max():.2e}")
        assert err.mean() <= 2e-3 and q <= 2e-2, (err.max(), q, err.mean())


def test_bf16_psnr_gate_structured_target(tn):
    """North star: BF16 within 0.1 dB PSNR of the fp32 path.  Scored against a structured target (the fp32 render of a
    DIFFERENT parameter set from the same view), 100x100, so a gross bf16 error moves the PSNR; the direct bf16-vs-fp32
    image PSNR is bounded as well."""
    from torch_nerf_b200.engine import HotPathEngine

    h = w = 100
    focal = orc.blender_focal(w)
    cam = tn.PerspectiveCamera({"f_x": focal, "f_y": focal, "img_width": w, "img_height": h},
                               torch.from_numpy(orc.pose_spherical(45.0, -30.0, 4.0)), 2.0, 6.0)
    n = h * w
    gen = torch.Generator(device="cuda").manual_seed(7)
    u = tuple(torch.rand((n, k), device="cuda", generator=gen) for k in (SC, SC, SF, SF))

    def render(seed_c, seed_f, precision):
        (c, f), _ = make_nets(tn, seed_c, seed_f, precision)
        eng = HotPathEngine(c, f, SC, SF, precision=precision)
        return eng.render_frame(cam, False, 0, n, max_rays=(1 << 20) if precision == "bf16" else 4096, uniforms=u).clone()

    target = render(61, 62, "fp32")
    a = render(51, 52, "fp32")
    b = render(51, 52, "bf16")

    def psnr(x, y):
        return float(-10.0 * torch.log10(torch.mean((x - y) ** 2)))

    p32, p16, direct = psnr(a, target), psnr(b, target), psnr(a, b)
    print(f"PSNR vs structured target: fp32 {p32:.3f} dB, bf16 {p16:.3f} dB, delta {p16 - p32:+.4f} dB; bf16 vs fp32 {direct:.1f} dB")
    assert float((a - target).abs().mean()) > 5e-3, "the target must differ from the render for the gate to mean anything"
    assert abs(p32 - p16) < 0.1, (p32, p16)
    assert direct > 35.0, direct
