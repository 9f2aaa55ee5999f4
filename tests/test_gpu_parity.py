"""GPU parity tests (run on the B200 box with `-m gpu`): the CUDA path, called through the C-ABI library via the
Python mirror classes, against the golden vectors (reference outputs) and the numpy oracle on identical inputs
and identical uniforms.

Gates (BASELINE.json north_star): fine-sample bin indices bit-exact; per-ray RGB / weights <= 1e-3 max-abs in
fp32-validation mode."""
import numpy as np
import pytest
import torch

from conftest import check_digest, load_golden
from oracle import nerf_oracle as orc

pytestmark = pytest.mark.gpu

TOL_FP32 = 1e-3  # north-star tolerance for fp32-validation mode (max abs)


@pytest.fixture(scope="module")
def tn():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import torch_nerf_b200 as mod

    mod._lib.load()
    return mod


def cu(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a)).to("cuda", dtype)


def blender_camera(tn, h, w, focal, c2w, near=2.0, far=6.0):
    return tn.PerspectiveCamera({"f_x": focal, "f_y": focal, "img_width": w, "img_height": h}, torch.from_numpy(c2w), near, far)


def load_params(net, params):
    net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in params.items()})
    return net.cuda()


def make_scene(tn, seed, precision="fp32"):
    params = orc.init_nerf_params(seed=seed)
    net = load_params(tn.NeRF(63, 27, precision=precision), params)
    enc = {"coord_enc": tn.PositionalEncoder(3, 10, True), "dir_enc": tn.PositionalEncoder(3, 4, True)}
    return tn.PrimitiveCube(net, enc), net, params


# ------------------------------------------------------------------------------------------------ K1
def test_raygen_golden(tn):
    g = load_golden("raygen.npz")
    sampler = tn.StratifiedSampler()
    h, w, focal = int(g["h"]), int(g["w"]), float(g["focal"])
    cam = blender_camera(tn, h, w, focal, g["c2w"])
    b = sampler.generate_rays(torch.from_numpy(g["coords"]), cam, project_to_ndc=False)
    assert np.array_equal(b.ray_origin.cpu().numpy(), g["ray_o"])
    np.testing.assert_allclose(b.ray_dir.cpu().numpy(), g["ray_d"], rtol=1e-6, atol=1e-7)
    b2 = sampler.generate_rays_from_pixels(torch.from_numpy(g["pix"]), cam, False)
    assert torch.equal(b2.ray_dir, b.ray_dir) and torch.equal(b2.ray_origin, b.ray_origin)
    ren = tn.VolumeRenderer(tn.QuadratureIntegrator(), sampler, cam)
    assert np.array_equal(ren.screen_coords[torch.from_numpy(g["pix"])].numpy(), g["coords"])
    h2, w2, f2 = int(g["h2"]), int(g["w2"]), float(g["focal2"])
    for tag, near in (("ndc0", 0.0), ("ndc1", 1.0)):
        cam2 = blender_camera(tn, h2, w2, f2, g["c2w2"], near, 1.0)
        b = sampler.generate_rays(torch.from_numpy(g[f"{tag}_coords"]), cam2, project_to_ndc=True)
        np.testing.assert_allclose(b.ray_origin.cpu().numpy(), g[f"{tag}_o"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(b.ray_dir.cpu().numpy(), g[f"{tag}_d"], rtol=2e-5, atol=1e-6)
        assert b.is_ndc and b.t_near == near


def test_raygen_full_frame_matches_oracle(tn):
    h, w = 100, 100
    focal = orc.blender_focal(w)
    c2w = orc.pose_spherical(12.0, -30.0, 4.0)
    cam = blender_camera(tn, h, w, focal, c2w)
    b = tn.StratifiedSampler().generate_rays_from_pixels(None, cam, False, first_pixel=0, count=h * w)
    o, d = orc.generate_rays(orc.screen_coords(h, w), orc.make_intrinsic(focal, focal, w, h), c2w, 2.0, h, w, False)
    np.testing.assert_allclose(b.ray_dir.cpu().numpy(), d, rtol=1e-6, atol=1e-7)
    assert np.array_equal(b.ray_origin.cpu().numpy(), o)


def test_raygen_frame_sized_launch_matches_oracle(tn):
    """>= 2^18 rays take the four-rays-per-thread kernel with 16-byte stores (smaller launches: one ray per thread); also an
    odd count and an unaligned output view, which must fall back to scalar stores."""
    h, w = 601, 603
    focal = orc.blender_focal(w)
    c2w = orc.pose_spherical(-71.0, -25.0, 4.0)
    cam = blender_camera(tn, h, w, focal, c2w)
    o, d = orc.generate_rays(orc.screen_coords(h, w), orc.make_intrinsic(focal, focal, w, h), c2w, 2.0, h, w, False)
    b = tn.StratifiedSampler().generate_rays_from_pixels(None, cam, False, first_pixel=0, count=h * w)
    np.testing.assert_allclose(b.ray_dir.cpu().numpy(), d, rtol=1e-6, atol=1e-7)
    assert np.array_equal(b.ray_origin.cpu().numpy(), o)
    lib = tn._lib.load()
    n = h * w
    buf_o = torch.zeros(3 * n + 1, device="cuda")
    buf_d = torch.zeros(3 * n + 1, device="cuda")
    tn._lib.check(lib.nerf_generate_rays_from_pixels(None, 0, n, cam.pack(False), tn._lib.c_void_p(buf_o.data_ptr() + 4),
                                                     tn._lib.c_void_p(buf_d.data_ptr() + 4), tn._lib.stream()), "raygen")
    assert torch.equal(buf_d[1:].view(n, 3), b.ray_dir) and torch.equal(buf_o[1:].view(n, 3), b.ray_origin)


# ------------------------------------------------------------------------------------------------ K2 / K3
def test_coarse_sampling_bit_exact(tn):
    g = load_golden("coarse.npz")
    sampler = tn.StratifiedSampler()
    for tag, (near, far) in (("b", (2.0, 6.0)), ("n", (0.0, 1.0))):
        bundle = tn.RayBundle(cu(g["ray_o"]), cu(g["ray_d"]), near, far, False)
        pts, dirs, delta = sampler.sample_along_rays(bundle, 64, device=0, uniforms=(cu(g["u"]),))
        assert np.array_equal(delta.cpu().numpy(), g[f"{tag}_delta"])
        assert np.array_equal(dirs.cpu().numpy(), g[f"{tag}_dirs"])
        assert np.array_equal(pts.cpu().numpy(), g[f"{tag}_pts"])


def test_fine_sampling_bit_exact(tn):
    g = load_golden("fine.npz")
    sampler = tn.StratifiedSampler()
    for tag, (near, far) in (("b", (2.0, 6.0)), ("n", (0.0, 1.0))):
        bundle = tn.RayBundle(cu(g["ray_o"]), cu(g["ray_d"]), near, far, False)
        w = cu(g["weights"])
        pts, dirs, delta, ex = sampler.sample_along_rays(
            bundle, (64, 128), device=0, weights=w, uniforms=(cu(g["u0"]), cu(g["u1"]), cu(g["u2"])), return_extras=True
        )
        assert np.array_equal(ex["idx"].cpu().numpy(), g[f"{tag}_idx"]), "bin indices must be bit-exact"
        assert np.array_equal(w.cpu().numpy(), g[f"{tag}_w_after"]), "weights += 1e-5 happens in place"
        assert np.array_equal(delta.cpu().numpy(), g[f"{tag}_delta"])
        assert np.array_equal(pts.cpu().numpy(), g[f"{tag}_pts"])
        t = ex["t"].cpu().numpy()
        assert np.all(np.diff(t, axis=-1) >= 0), "samples must come out sorted"
        # sample_pdf on its own
        bins, step = tn.make_bins(near, far, 64)
        t_f, idx = tn.sample_pdf(bins[None].repeat(w.shape[0], 1), step, cu(g["weights"]), 128,
                                 uniforms=(cu(g["u1"]), cu(g["u2"])), return_indices=True)
        assert np.array_equal(idx.cpu().numpy(), g[f"{tag}_idx"])
        assert np.array_equal(t_f.cpu().numpy(), g[f"{tag}_t_fine"])


def test_fine_idx_big_bit_exact(tn):
    g = load_golden("fine_idx_big.npz")
    rng = np.random.default_rng(int(g["seed"]))
    n, sc, sf = 4096, 64, 128
    w = (rng.random((n, sc), dtype=np.float32) ** 4 * rng.random((n, 1), dtype=np.float32)).astype(np.float32)
    u1 = rng.random((n, sf), dtype=np.float32)
    bins, step = tn.make_bins(2.0, 6.0, sc)
    _, idx = tn.sample_pdf(bins[None].repeat(n, 1), step, cu(w), sf, uniforms=(cu(u1), cu(u1)), return_indices=True)
    assert np.array_equal(idx.cpu().numpy().astype(np.uint8), g["idx"])


def test_fine_sampling_properties_full_size(tn):
    """C2-sized (4096 rays, 64+128): sortedness, every importance sample lies inside its bin, deltas consistent."""
    torch.manual_seed(0)
    n, sc, sf = 4096, 64, 128
    ray_o = torch.randn(n, 3, device="cuda")
    ray_d = torch.randn(n, 3, device="cuda")
    w = torch.rand(n, sc, device="cuda") ** 8
    bundle = tn.RayBundle(ray_o, ray_d, 2.0, 6.0, False)
    pts, dirs, delta, ex = tn.StratifiedSampler().sample_along_rays(bundle, (sc, sf), device=0, weights=w, return_extras=True)
    t = ex["t"]
    assert torch.all(t[:, 1:] >= t[:, :-1]) and t.min() >= 2.0 and t.max() < 6.0 + 1e-5
    assert torch.equal(delta[:, :-1], t[:, 1:] - t[:, :-1])
    assert torch.equal(delta[:, -1], torch.full((n,), 1e8, device="cuda") - t[:, -1])
    assert ex["idx"].min() >= 0 and ex["idx"].max() <= sc - 1
    assert torch.equal(dirs, ray_d[:, None, :].expand(n, sc + sf, 3))
    assert torch.equal(pts, ray_o[:, None, :] + t[..., None] * ray_d[:, None, :])


@pytest.mark.parametrize("s,near,far", [(64, 2.0, 6.0), (192, 0.0, 1.0), (60, 0.37, 5.13), (7, 2.0, 6.0)])
def test_coarse_fused_form_equals_materialised_form(tn, s, near, far):
    """The elementwise kernel behind materialize=False (samples % 4 == 0) and the warp-per-ray kernel must agree bit for
    bit with each other and, for float32-exact bins, with the oracle."""
    rng = np.random.default_rng(s)
    n = 5003
    o, d = rng.normal(size=(n, 3)).astype(np.float32), rng.normal(size=(n, 3)).astype(np.float32)
    u = rng.random((n, s), dtype=np.float32)
    bundle = tn.RayBundle(cu(o), cu(d), near, far, False)
    sampler = tn.StratifiedSampler()
    _, _, delta_m, ex_m = sampler.sample_along_rays(bundle, s, device=0, uniforms=(cu(u),), return_extras=True)
    _, _, delta_f, ex_f = sampler.sample_along_rays(bundle, s, device=0, uniforms=(cu(u),), return_extras=True, materialize=False)
    assert torch.equal(ex_m["t"], ex_f["t"]) and torch.equal(delta_m, delta_f)
    _, _, delta_o, t_o = orc.sample_along_rays_coarse(o, d, near, far, s, u)
    assert np.array_equal(ex_f["t"].cpu().numpy(), t_o) and np.array_equal(delta_f.cpu().numpy(), delta_o)


@pytest.mark.parametrize("sc,sf", [(64, 128), (32, 64), (64, 64), (96, 128)])
@pytest.mark.parametrize("kind", ["peaked", "zeros", "wide"])
def test_fine_sampling_paths_vs_oracle(tn, sc, sf, kind):
    """Every fine-sampling path -- the register-sort kernel of the default 64+128, the general shared-memory sort, the
    warp-scan CDF and its sequential fall-back (`wide`: pdf values below 2^-28) -- against the oracle, bit for bit."""
    rng = np.random.default_rng(sc * 1000 + sf)
    n = 2051
    o, d = rng.normal(size=(n, 3)).astype(np.float32), rng.normal(size=(n, 3)).astype(np.float32)
    if kind == "peaked":
        w = (rng.random((n, sc), dtype=np.float32) ** 8).astype(np.float32)
    elif kind == "zeros":
        w = np.zeros((n, sc), np.float32)
        w[::2, sc // 3] = 0.9
    else:
        w = np.where(rng.random((n, sc)) < 0.1, 3e4, 1e-7).astype(np.float32)
    u0, u1, u2 = (rng.random((n, k), dtype=np.float32) for k in (sc, sf, sf))
    bundle = tn.RayBundle(cu(o), cu(d), 2.0, 6.0, False)
    w_dev = cu(w)
    pts, dirs, delta, ex = tn.StratifiedSampler().sample_along_rays(
        bundle, (sc, sf), device=0, weights=w_dev, uniforms=(cu(u0), cu(u1), cu(u2)), return_extras=True)
    w_o = w.copy()
    pts_o, _, delta_o, t_o, idx_o = orc.sample_along_rays_fine(o, d, 2.0, 6.0, sc, sf, w_o, u0, u1, u2)
    assert np.array_equal(ex["idx"].cpu().numpy(), idx_o), "bin indices must be bit-exact"
    assert np.array_equal(w_dev.cpu().numpy(), w_o)
    assert np.array_equal(ex["t"].cpu().numpy(), t_o)
    assert np.array_equal(delta.cpu().numpy(), delta_o)
    assert np.array_equal(pts.cpu().numpy(), pts_o)
    # fused form (no (N,S,3) outputs) gives the same t / delta
    w_dev2 = cu(w)
    _, _, delta2, ex2 = tn.StratifiedSampler().sample_along_rays(
        bundle, (sc, sf), device=0, weights=w_dev2, uniforms=(cu(u0), cu(u1), cu(u2)), return_extras=True, materialize=False)
    assert torch.equal(ex2["t"], ex["t"]) and torch.equal(delta2, delta)


@pytest.mark.parametrize("near,far", [(0.37, 5.13), (0.0, 1.0), (1.7, 93.1)])
def test_fine_sampling_inexact_bins_vs_oracle(tn, near, far):
    """Scene bounds whose bin edges are not exact in float32: the stratified draw may come out locally unordered by an
    ulp, which the 64+128 kernel must detect (it skips the coarse sort otherwise)."""
    rng = np.random.default_rng(77)
    n, sc, sf = 4099, 64, 128
    o, d = rng.normal(size=(n, 3)).astype(np.float32), rng.normal(size=(n, 3)).astype(np.float32)
    w = (rng.random((n, sc), dtype=np.float32) ** 6).astype(np.float32)
    u0, u1, u2 = (rng.random((n, k), dtype=np.float32) for k in (sc, sf, sf))
    u0[::3] = np.float32(1.0) - np.float32(2.0 ** -24)  # draws at the top of every bin: neighbours collide
    u0[1::3, ::2] = 0.0
    bundle = tn.RayBundle(cu(o), cu(d), near, far, False)
    w_dev = cu(w)
    _, _, delta, ex = tn.StratifiedSampler().sample_along_rays(
        bundle, (sc, sf), device=0, weights=w_dev, uniforms=(cu(u0), cu(u1), cu(u2)), return_extras=True, materialize=False)
    _, _, delta_o, t_o, idx_o = orc.sample_along_rays_fine(o, d, near, far, sc, sf, w.copy(), u0, u1, u2)
    assert np.array_equal(ex["idx"].cpu().numpy(), idx_o)
    assert np.array_equal(ex["t"].cpu().numpy(), t_o)
    assert np.array_equal(delta.cpu().numpy(), delta_o)


def test_sampler_value_errors(tn):
    bundle = tn.RayBundle(torch.zeros(4, 3, device="cuda"), torch.ones(4, 3, device="cuda"), 2.0, 6.0, False)
    s = tn.StratifiedSampler()
    with pytest.raises(ValueError):
        s.sample_along_rays(bundle, (64, 128), device=0)  # tuple without weights
    with pytest.raises(ValueError):
        s.sample_along_rays(bundle, 64, device=0, weights=torch.ones(4, 64, device="cuda"))
    with pytest.raises(ValueError):
        s.sample_along_rays(bundle, (64, 128), device=0, weights=[1.0])
    # empty ray set is a no-op
    empty = tn.RayBundle(torch.zeros(0, 3, device="cuda"), torch.zeros(0, 3, device="cuda"), 2.0, 6.0, False)
    pts, dirs, delta = s.sample_along_rays(empty, 64, device=0)
    assert pts.shape == (0, 64, 3) and delta.shape == (0, 64)


# ------------------------------------------------------------------------------------------------ K4
def test_posenc_golden(tn):
    g = load_golden("posenc.npz")
    x = cu(g["x"])
    # CUDA sincosf: <= 2 ulp of the result over the full range
    for enc, key in ((tn.PositionalEncoder(3, 10, True), "out10"), (tn.PositionalEncoder(3, 4, True), "out4"),
                     (tn.PositionalEncoder(3, 4, False), "out4_noinput")):
        out = enc.encode(x).cpu().numpy()
        assert out.shape == g[key].shape and enc.out_dim == g[key].shape[1]
        np.testing.assert_allclose(out, g[key], rtol=0, atol=5e-7)
    big = torch.randn(100_003, 3, device="cuda") * 3
    out = tn.PositionalEncoder(3, 10, True).encode(big)
    ref = orc.positional_encode(big.cpu().numpy(), 10)
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=0, atol=5e-7)


# ------------------------------------------------------------------------------------------------ K7 / K8
def test_composite_golden(tn):
    g = load_golden("composite.npz")
    integ = tn.QuadratureIntegrator()
    sigma = cu(g["sigma"]).requires_grad_(True)
    rad = cu(g["radiance"]).requires_grad_(True)
    rgb, w = integ.integrate_along_rays(sigma, rad, cu(g["delta"]))
    np.testing.assert_allclose(w.detach().cpu().numpy(), g["w"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(rgb.detach().cpu().numpy(), g["rgb"], rtol=1e-5, atol=2e-6)
    (rgb * cu(g["g_rgb"])).sum().backward(retain_graph=True)
    np.testing.assert_allclose(rad.grad.cpu().numpy(), g["g_radiance"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(sigma.grad.cpu().numpy(), g["g_sigma"], rtol=5e-4, atol=5e-5)
    sigma.grad = None
    rad.grad = None
    ((rgb * cu(g["g_rgb"])).sum() + (w * cu(g["g_w"])).sum()).backward()
    np.testing.assert_allclose(rad.grad.cpu().numpy(), g["g_radiance_w"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(sigma.grad.cpu().numpy(), g["g_sigma_w"], rtol=5e-4, atol=5e-5)


@pytest.mark.parametrize("s", [64, 192, 100, 33, 256, 300])
def test_composite_full_size_vs_oracle(tn, s):
    rng = np.random.default_rng(10 + s)
    n = 4096
    sigma = np.maximum(rng.normal(size=(n, s)) * 2.0, 0).astype(np.float32)
    rad = rng.random((n, s, 3), dtype=np.float32)
    t = np.sort(2.0 + 4.0 * rng.random((n, s)), axis=-1).astype(np.float32)
    delta = np.diff(np.concatenate([t, np.full((n, 1), 1e8, np.float32)], -1), axis=-1).astype(np.float32)
    g_rgb = rng.normal(size=(n, 3)).astype(np.float32)
    rgb_o, w_o = orc.integrate_along_rays(sigma, rad, delta)
    gs_o, gc_o = orc.integrate_along_rays_backward(sigma, rad, delta, g_rgb)
    sg, rd = cu(sigma).requires_grad_(True), cu(rad).requires_grad_(True)
    rgb, w = tn.QuadratureIntegrator().integrate_along_rays(sg, rd, cu(delta))
    np.testing.assert_allclose(w.detach().cpu().numpy(), w_o, rtol=0, atol=2e-6)
    np.testing.assert_allclose(rgb.detach().cpu().numpy(), rgb_o, rtol=0, atol=1e-5)
    # property: weights are a sub-probability vector; opacity + residual transmittance == 1
    assert float(w.sum(-1).max()) <= 1.0 + 1e-5
    (rgb * cu(g_rgb)).sum().backward()
    np.testing.assert_allclose(rd.grad.cpu().numpy(), gc_o, rtol=1e-5, atol=2e-6)
    scale = np.abs(gs_o).max()
    np.testing.assert_allclose(sg.grad.cpu().numpy() / scale, gs_o / scale, rtol=0, atol=2e-6)
    depth = tn.QuadratureIntegrator().integrate_with_depth(cu(sigma), cu(rad), cu(delta), cu(t))
    np.testing.assert_allclose(depth[2].cpu().numpy(), (w_o * t).sum(-1), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(depth[3].cpu().numpy(), w_o.sum(-1), rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------------------------------------ K5 / K6 (fp32)
def test_mlp_f32_golden(tn):
    g = load_golden("mlp.npz")
    params = orc.init_nerf_params(seed=int(g["seed"]))
    net = load_params(tn.NeRF(63, 27), params)
    sigma, rgb = net(cu(g["pe"]), cu(g["de"]))
    np.testing.assert_allclose(sigma.detach().cpu().numpy(), g["sigma"], rtol=1e-4, atol=5e-6)
    np.testing.assert_allclose(rgb.detach().cpu().numpy(), g["rgb"], rtol=1e-4, atol=5e-6)
    ((sigma * cu(g["g_sigma"])).sum() + (rgb * cu(g["g_rgb"])).sum()).backward()
    grads = {k: p.grad.cpu().numpy() for k, p in net.named_parameters()}
    check_digest(grads, g, rtol=5e-4, atol=5e-6)
    with pytest.raises(ValueError):
        net(cu(g["pe"])[None], cu(g["de"]))
    with pytest.raises(ValueError):
        net(cu(g["pe"]), cu(g["de"])[:-1])
    with pytest.raises(ValueError):
        net(cu(g["pe"])[:, :-1], cu(g["de"]))


def test_mlp_f32_vs_oracle_odd_sizes(tn):
    """Ragged row counts (not multiples of any tile) and a non-default architecture."""
    rng = np.random.default_rng(5)
    for (p, v, f, m) in ((63, 27, 256, 1000), (39, 15, 64, 77), (63, 27, 256, 1)):
        params = orc.init_nerf_params(p, v, f, seed=3)
        net = load_params(tn.NeRF(p, v, f), params)
        pos = rng.normal(size=(m, p)).astype(np.float32)
        view = rng.normal(size=(m, v)).astype(np.float32)
        g_s = rng.normal(size=(m,)).astype(np.float32)
        g_c = rng.normal(size=(m, 3)).astype(np.float32)
        s_o, c_o, acts = orc.nerf_forward(params, pos, view, return_cache=True)
        grads_o = orc.nerf_backward(params, acts, g_s, g_c)
        sigma, rgb = net(cu(pos), cu(view))
        np.testing.assert_allclose(sigma.detach().cpu().numpy(), s_o, rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(rgb.detach().cpu().numpy(), c_o, rtol=1e-4, atol=1e-5)
        ((sigma * cu(g_s)).sum() + (rgb * cu(g_c)).sum()).backward()
        for k, prm in net.named_parameters():
            ref = grads_o[k]
            np.testing.assert_allclose(prm.grad.cpu().numpy(), ref, rtol=2e-3, atol=2e-4 * (np.abs(ref).max() + 1e-6), err_msg=k)


# ------------------------------------------------------------------------------------------------ L1 render + train step
def test_render_scene_golden_fp32(tn):
    """Config C1 in miniature through VolumeRenderer.render_scene (coarse then fine, 2 ray batches)."""
    g = load_golden("render.npz")
    h, w, focal = int(g["h"]), int(g["w"]), float(g["focal"])
    cam = blender_camera(tn, h, w, focal, g["c2w"])
    scene_c, _, _ = make_scene(tn, int(g["seed_c"]))
    scene_f, _, _ = make_scene(tn, int(g["seed_f"]))
    ren = tn.VolumeRenderer(tn.QuadratureIntegrator(), tn.StratifiedSampler(), cam)
    with torch.no_grad():
        rgb_c, pix, w_c = ren.render_scene(scene_c, h * w, 64, False, 0, num_ray_batch=2, uniforms=(cu(g["u_c"]),))
        assert np.array_equal(pix.numpy(), g["pix"])
        np.testing.assert_allclose(rgb_c.cpu().numpy(), g["rgb_c"], rtol=0, atol=TOL_FP32)
        np.testing.assert_allclose(w_c.cpu().numpy(), g["w_c"], rtol=0, atol=TOL_FP32)
        # fine pass from the reference's coarse weights: bin decisions on identical inputs
        rgb_f, _, w_f = ren.render_scene(scene_f, h * w, (64, 128), False, 0, pixel_indices=pix, weights=cu(g["w_c"]),
                                         num_ray_batch=2, uniforms=(cu(g["u0"]), cu(g["u1"]), cu(g["u2"])))
    np.testing.assert_allclose(rgb_f.cpu().numpy(), g["rgb_f"], rtol=0, atol=TOL_FP32)
    np.testing.assert_allclose(w_f.cpu().numpy(), g["w_f"], rtol=0, atol=TOL_FP32)
    # tighter than the gate in practice
    assert np.abs(rgb_f.cpu().numpy() - g["rgb_f"]).max() < 5e-5
    with pytest.raises(ValueError):
        ren.render_scene(scene_c, 10.0, 64, False, 0)
    with pytest.raises(ValueError):
        ren.render_scene(scene_c, 10, (64, 128, 1), False, 0, pixel_indices=pix)
    with pytest.raises(ValueError):
        ren.render_scene(scene_c, 10, (64, 128), False, 0)


def test_train_step_golden_fp32(tn):
    """Config C2 in miniature: train.py:130-218 up to backward, losses and all 44 gradient tensors."""
    g = load_golden("train_step.npz")
    h, w, focal = int(g["h"]), int(g["w"]), float(g["focal"])
    cam = blender_camera(tn, h, w, focal, g["c2w"])
    scene_c, net_c, _ = make_scene(tn, int(g["seed_c"]))
    scene_f, net_f, _ = make_scene(tn, int(g["seed_f"]))
    ren = tn.VolumeRenderer(tn.QuadratureIntegrator(), tn.StratifiedSampler(), cam)
    pix = torch.from_numpy(g["pix"])
    n = pix.shape[0]
    target = cu(g["target"])
    loss_fn = torch.nn.MSELoss()
    pred_c, idx_c, w_c = ren.render_scene(scene_c, n, 64, False, 0, pixel_indices=pix, uniforms=(cu(g["u_c"]),))
    loss_c = loss_fn(target, pred_c)
    pred_f, _, w_f = ren.render_scene(scene_f, n, (64, 128), False, 0, pixel_indices=idx_c, weights=w_c,
                                      uniforms=(cu(g["u0"]), cu(g["u1"]), cu(g["u2"])))
    loss_f = loss_fn(target, pred_f)
    (loss_c + loss_f).backward()
    np.testing.assert_allclose(pred_c.detach().cpu().numpy(), g["rgb_c"], rtol=0, atol=TOL_FP32)
    np.testing.assert_allclose(pred_f.detach().cpu().numpy(), g["rgb_f"], rtol=0, atol=TOL_FP32)
    np.testing.assert_allclose(loss_c.item(), float(g["loss_c"]), rtol=1e-4)
    np.testing.assert_allclose(loss_f.item(), float(g["loss_f"]), rtol=1e-3)
    check_digest({k: p.grad.cpu().numpy() for k, p in net_c.named_parameters()}, g, prefix="c/", rtol=5e-3, atol=1e-6)
    check_digest({k: p.grad.cpu().numpy() for k, p in net_f.named_parameters()}, g, prefix="f/", rtol=5e-3, atol=1e-6)


@pytest.mark.parametrize("near", [0.0, 1.0])
def test_map_rays_to_ndc_standalone(tn, near):
    """RaySamplerBase.map_rays_to_ndc (sampler_base.py:199-257) on its own against the oracle's restatement."""
    rng = np.random.default_rng(3)
    n = 5000
    o = rng.normal(size=(n, 3)).astype(np.float32)
    o[:, 2] = np.abs(o[:, 2]) + 0.5
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d[:, 2] = -np.abs(d[:, 2]) - 0.1
    ro, rd = orc.map_rays_to_ndc(815.0, near, 756, 1008, o, d)
    po, pd = tn.StratifiedSampler().map_rays_to_ndc(815.0, near, 756, 1008, torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda())
    np.testing.assert_allclose(po.cpu().numpy(), ro, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(pd.cpu().numpy(), rd, rtol=1e-6, atol=1e-6)
    with pytest.raises(ValueError):
        tn.StratifiedSampler().map_rays_to_ndc(815.0, -1.0, 756, 1008, torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda())
