"""CPU: pins the numpy oracle (oracle/nerf_oracle.py) against the golden vectors produced by the
unmodified reference (tests/golden/make_golden.py)."""

import numpy as np
import pytest

from conftest import check_digest, load_golden
from oracle import nerf_oracle as orc


def test_screen_coords_and_partitions():
    c = orc.screen_coords(3, 4)
    assert c.dtype == np.int64 and c.shape == (12, 2)
    assert c[0].tolist() == [0, 2] and c[5].tolist() == [1, 1] and c[11].tolist() == [3, 0]
    # volume_renderer.py:229-235: 800x800 / 4096 -> 156 chunks; 100x100 -> 2 chunks of 5000
    p = orc.ray_batch_partitions(640000, 156)
    assert p[0] == 0 and p[-1] == 640000 and len(p) == 157 and np.all(np.diff(p) > 4000)
    assert orc.ray_batch_partitions(10000, 2).tolist() == [0, 5000, 10000]


def test_raygen_golden():
    g = load_golden("raygen.npz")
    h, w, focal = int(g["h"]), int(g["w"]), float(g["focal"])
    coords = orc.screen_coords(h, w)[g["pix"]]
    assert np.array_equal(coords, g["coords"])
    o, d = orc.generate_rays(coords, orc.make_intrinsic(focal, focal, w, h), g["c2w"], 2.0, h, w, False)
    np.testing.assert_allclose(o, g["ray_o"], rtol=0, atol=0)
    np.testing.assert_allclose(d, g["ray_d"], rtol=1e-6, atol=1e-7)
    h2, w2, f2 = int(g["h2"]), int(g["w2"]), float(g["focal2"])
    coords2 = orc.screen_coords(h2, w2)[g["pix2"]]
    for tag, near in (("ndc0", 0.0), ("ndc1", 1.0)):
        assert np.array_equal(coords2, g[f"{tag}_coords"])
        o, d = orc.generate_rays(coords2, orc.make_intrinsic(f2, f2, w2, h2), g["c2w2"], near, h2, w2, True)
        np.testing.assert_allclose(o, g[f"{tag}_o"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(d, g[f"{tag}_d"], rtol=2e-5, atol=1e-6)


def test_bins_match_torch_linspace():
    torch = pytest.importorskip("torch")
    for near, far, s in ((2.0, 6.0, 64), (0.0, 1.0, 64), (2.0, 6.0, 128), (0.5, 3.5, 32)):
        bins, step = orc.create_t_bins(near, far, s)
        ref = torch.linspace(near, far, s + 1)[:-1].numpy()
        assert np.array_equal(bins, ref), (near, far, s)
        assert step == (far - near) / s


def test_coarse_golden():
    g = load_golden("coarse.npz")
    for tag, (near, far) in (("b", (2.0, 6.0)), ("n", (0.0, 1.0))):
        pts, dirs, delta, t = orc.sample_along_rays_coarse(g["ray_o"], g["ray_d"], near, far, 64, g["u"])
        assert np.array_equal(delta, g[f"{tag}_delta"])
        assert np.array_equal(dirs, g[f"{tag}_dirs"])
        assert np.array_equal(pts, g[f"{tag}_pts"])


def test_fine_golden_bit_exact():
    g = load_golden("fine.npz")
    for tag, (near, far) in (("b", (2.0, 6.0)), ("n", (0.0, 1.0))):
        w = g["weights"].copy()
        pts, dirs, delta, t, idx = orc.sample_along_rays_fine(g["ray_o"], g["ray_d"], near, far, 64, 128, w,
                                                              g["u0"], g["u1"], g["u2"])
        assert np.array_equal(idx, g[f"{tag}_idx"]), "bin indices must be bit-exact"
        assert np.array_equal(w, g[f"{tag}_w_after"]), "in-place += 1e-5 side effect"
        assert np.array_equal(delta, g[f"{tag}_delta"])
        assert np.array_equal(pts, g[f"{tag}_pts"])
        bins, step = orc.create_t_bins(near, far, 64)
        t_f, idx2 = orc.sample_pdf(np.repeat(bins[None], w.shape[0], 0), step, g["weights"].copy(), g["u1"], g["u2"])
        assert np.array_equal(t_f, g[f"{tag}_t_fine"]) and np.array_equal(idx2, idx)


def test_fine_idx_big_bit_exact():
    g = load_golden("fine_idx_big.npz")
    rng = np.random.default_rng(int(g["seed"]))
    n, sc, sf = 4096, 64, 128
    w = (rng.random((n, sc), dtype=np.float32) ** 4 * rng.random((n, 1), dtype=np.float32)).astype(np.float32)
    u1 = rng.random((n, sf), dtype=np.float32)
    cdf = orc.pdf_to_cdf(w)
    assert np.array_equal(cdf[:, -1], g["cdf_last"])
    idx = (cdf[:, None, :] <= u1[:, :, None]).sum(-1) - 1
    assert np.array_equal(idx.astype(np.uint8), g["idx"])


def test_posenc_golden():
    g = load_golden("posenc.npz")
    # sin/cos of arguments up to 2^9 * 6: numpy vs torch(SLEEF) agree to ~1 ulp of the result
    np.testing.assert_allclose(orc.positional_encode(g["x"], 10), g["out10"], rtol=0, atol=3e-7)
    np.testing.assert_allclose(orc.positional_encode(g["x"], 4), g["out4"], rtol=0, atol=3e-7)
    np.testing.assert_allclose(orc.positional_encode(g["x"], 4, False), g["out4_noinput"], rtol=0, atol=3e-7)
    assert orc.positional_encode(g["x"], 10).shape[1] == 63


def test_mlp_golden():
    g = load_golden("mlp.npz")
    params = orc.init_nerf_params(seed=int(g["seed"]))
    pe = orc.positional_encode(g["pts"], 10)
    de = orc.positional_encode(g["dirs"], 4)
    sigma, rgb, acts = orc.nerf_forward(params, pe, de, return_cache=True)
    np.testing.assert_allclose(sigma, g["sigma"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(rgb, g["rgb"], rtol=1e-4, atol=2e-6)
    grads = orc.nerf_backward(params, acts, g["g_sigma"], g["g_rgb"])
    check_digest(grads, g)
    with pytest.raises(ValueError):
        orc.nerf_forward(params, pe[None], de)
    with pytest.raises(ValueError):
        orc.nerf_forward(params, pe, de[:-1])


def test_composite_golden():
    g = load_golden("composite.npz")
    rgb, w = orc.integrate_along_rays(g["sigma"], g["radiance"], g["delta"])
    np.testing.assert_allclose(w, g["w"], rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(rgb, g["rgb"], rtol=1e-5, atol=1e-6)
    gs, gc = orc.integrate_along_rays_backward(g["sigma"], g["radiance"], g["delta"], g["g_rgb"])
    np.testing.assert_allclose(gc, g["g_radiance"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(gs, g["g_sigma"], rtol=2e-4, atol=2e-5)
    gs, gc = orc.integrate_along_rays_backward(g["sigma"], g["radiance"], g["delta"], g["g_rgb"], g["g_w"])
    np.testing.assert_allclose(gc, g["g_radiance_w"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(gs, g["g_sigma_w"], rtol=2e-4, atol=2e-5)


def test_render_golden():
    g = load_golden("render.npz")
    h, w, focal = int(g["h"]), int(g["w"]), float(g["focal"])
    coords = orc.screen_coords(h, w)[g["pix"]]
    o, d = orc.generate_rays(coords, orc.make_intrinsic(focal, focal, w, h), g["c2w"], 2.0, h, w, False)
    pc = orc.init_nerf_params(seed=int(g["seed_c"]))
    pf = orc.init_nerf_params(seed=int(g["seed_f"]))
    co = orc.render_pass(pc, o, d, 2.0, 6.0, 64, (g["u_c"],), num_ray_batch=2)
    np.testing.assert_allclose(co["rgb"], g["rgb_c"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(co["weights"], g["w_c"], rtol=0, atol=1e-5)
    # fine pass from the REFERENCE's coarse weights so bin decisions are made on identical inputs
    fi = orc.render_pass(pf, o, d, 2.0, 6.0, (64, 128), (g["u0"], g["u1"], g["u2"]), weights=g["w_c"].copy(),
                         num_ray_batch=2)
    np.testing.assert_allclose(fi["rgb"], g["rgb_f"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(fi["weights"], g["w_f"], rtol=0, atol=1e-5)


def test_train_step_golden():
    g = load_golden("train_step.npz")
    h, w, focal = int(g["h"]), int(g["w"]), float(g["focal"])
    coords = orc.screen_coords(h, w)[g["pix"]]
    o, d = orc.generate_rays(coords, orc.make_intrinsic(focal, focal, w, h), g["c2w"], 2.0, h, w, False)
    pc = orc.init_nerf_params(seed=int(g["seed_c"]))
    pf = orc.init_nerf_params(seed=int(g["seed_f"]))
    out = orc.train_step_grads(pc, pf, o, d, 2.0, 6.0, 64, 128, g["target"], g["u_c"], g["u0"], g["u1"], g["u2"])
    np.testing.assert_allclose(out["coarse_rgb"], g["rgb_c"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(out["fine_rgb"], g["rgb_f"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(out["coarse_loss"], float(g["loss_c"]), rtol=1e-5)
    np.testing.assert_allclose(out["fine_loss"], float(g["loss_f"]), rtol=1e-4)
    check_digest(out["coarse_grads"], g, prefix="c/", rtol=2e-3, atol=1e-6)
    check_digest(out["fine_grads"], g, prefix="f/", rtol=2e-3, atol=1e-6)


def test_cdf_warp_scan_is_exactly_the_sequential_sum():
    """rays.cu build_cdf replaces torch.cumsum's sequential float64 running sum by a warp scan when every pdf value is
    in [2^-28, 1].  The claim is that all float64 partial sums are then exact, so the association order cannot matter:
    replay the kernel's Hillis-Steele scan (32-wide chunks + carry) in numpy and demand bit equality with np.cumsum."""
    rng = np.random.default_rng(5)
    for sc in (64, 96, 256):
        w = (rng.random((4096, sc), dtype=np.float32) ** 8).astype(np.float32)
        w[::7] = 0.0
        w += np.float32(1e-5)
        pdf = (w / orc.torch_cpu_sum_lastdim(w)[:, None]).astype(np.float32)
        assert pdf.min() >= 2.0 ** -28 and pdf.max() <= 1.0
        seq = np.cumsum(pdf.astype(np.float64), axis=-1)
        scan = np.empty_like(seq)
        carry = np.zeros(pdf.shape[0])
        for base in range(0, sc, 32):
            inc = pdf[:, base:base + 32].astype(np.float64)
            d = 1
            while d < 32:
                up = np.zeros_like(inc)
                up[:, d:] = inc[:, :-d]
                inc = inc + up
                d *= 2
            inc = inc + carry[:, None]
            scan[:, base:base + 32] = inc
            carry = inc[:, -1]
        assert np.array_equal(scan, seq)


def test_blocked_sorting_network_of_the_fine_sampler():
    """rays.cu warp_sort128_blocked, replayed in numpy stage by stage (same partner / mirror / keep-min rules): it must
    sort any 128 values, duplicates included."""
    rng = np.random.default_rng(0)
    lanes = np.arange(32)

    def cx(v, a, b):
        lo, hi = np.minimum(v[:, a], v[:, b]), np.maximum(v[:, a], v[:, b])
        v[:, a], v[:, b] = lo, hi

    def network(v):
        v = v.copy()
        k = 2
        while k <= 128:
            j = k >> 1
            while j > 0:
                mirror = j == (k >> 1)
                if j >= 4:
                    lane_mask = ((k - 1) >> 2) if mirror else (j >> 2)
                    keep_min = (lanes & (j >> 2)) == 0
                    new = v.copy()
                    for r in range(4):
                        other = (v[:, r ^ 3] if mirror else v[:, r])[lanes ^ lane_mask]
                        new[:, r] = np.where((v[:, r] > other) == keep_min, other, v[:, r])
                    v = new
                elif j == 2:
                    for a, b in ([(0, 3), (1, 2)] if mirror else [(0, 2), (1, 3)]):
                        cx(v, a, b)
                else:
                    cx(v, 0, 1)
                    cx(v, 2, 3)
                j >>= 1
            k <<= 1
        return v

    for trial in range(300):
        x = rng.random((32, 4)).astype(np.float32)
        if trial % 3 == 0:
            x = np.round(x * 8) / 8
        assert np.array_equal(network(x).reshape(-1), np.sort(x.reshape(-1)))


def test_blocked_transmittance_scan_rounds_like_the_sequential_cumsum():
    """encode_composite.cu composite_bwd_blk_kernel sums x = sigma*delta in float64 inside a lane (G consecutive samples),
    scans the lane totals across the warp and rounds every prefix to float32.  torch's cumsum is a sequential float64
    sum rounded per element; the two associations differ by ~1e-16 relative before the float32 rounding, so the rounded
    prefixes must agree everywhere except for (at most) isolated last-bit ties."""
    rng = np.random.default_rng(9)
    n, g = 2048, 6
    s = 32 * g
    sigma = np.maximum(rng.normal(size=(n, s)) * 3.0, 0).astype(np.float32)
    t = np.sort(2.0 + 4.0 * rng.random((n, s)), axis=-1).astype(np.float32)
    delta = np.diff(np.concatenate([t, np.full((n, 1), 1e8, np.float32)], -1), axis=-1).astype(np.float32)
    x = (sigma * delta).astype(np.float32)
    seq = np.cumsum(x.astype(np.float64), axis=-1).astype(np.float32)
    xb = x.astype(np.float64).reshape(n, 32, g)
    pre = np.cumsum(xb, axis=-1)                       # serial inside a lane
    tot = pre[:, :, -1].copy()
    inc = tot.copy()
    d = 1
    while d < 32:                                      # Hillis-Steele over the lane totals
        up = np.zeros_like(inc)
        up[:, d:] = inc[:, :-d]
        inc = inc + up
        d *= 2
    off = np.concatenate([np.zeros((n, 1)), inc[:, :-1]], axis=-1)
    blk = (off[:, :, None] + pre).reshape(n, s).astype(np.float32)
    mism = blk != seq
    assert mism.mean() < 1e-5
    if mism.any():
        assert np.all(np.abs(blk[mism] - seq[mism]) <= np.spacing(np.abs(seq[mism])))
