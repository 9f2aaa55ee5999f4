"""
CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

A numpy restatement of torch-NeRF's per-ray rendering hot path (ray generation ->
stratified / hierarchical sampling -> positional encoding -> 8x256 NeRF MLP -> alpha
compositing, forward and backward).  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module, and only as
the checker or the timed CPU baseline.  The product (`torch-nerf_b200/`) never imports it and
fails loudly when its CUDA library is missing.

Parity status: PINNED.  The reference ships no tests or golden vectors of its own
(SURVEY.md section 4), so the oracle is pinned against outputs of the unmodified reference
modules imported from /root/reference in the dev container; the generating script is
`tests/golden/make_golden.py` and the frozen vectors are `tests/golden/*.npz`
(`tests/test_oracle_golden.py` replays them on every CPU test run).

Every function cites the reference file:line it follows (paths relative to the reference
root, `torch_nerf/src/...`).  All arithmetic is float32 unless stated; places where the
reference's CPU kernels accumulate in a specific order (torch.sum) or in float64
(torch.cumsum) are restated explicitly because the fine-sample bin indices must be bit-exact.
"""

from __future__ import annotations

import numpy as np

F32 = np.float32

# --------------------------------------------------------------------------------------
# camera / screen  (renderer/cameras.py, renderer/volume_renderer.py)
# --------------------------------------------------------------------------------------


def make_intrinsic(fx: float, fy: float, img_w: float, img_h: float) -> np.ndarray:
    """renderer/cameras.py:84-118 -- 4x4 intrinsic with cx = W/2, cy = H/2."""
    return np.array(
        [[fx, 0.0, img_w / 2.0, 0.0], [0.0, fy, img_h / 2.0, 0.0], [0.0, 0.0, 0.0, 0.0], [0.0, 0.0, -1.0, 0.0]],
        dtype=F32,
    )


def screen_coords(img_h: int, img_w: int) -> np.ndarray:
    """renderer/volume_renderer.py:171-190 -- pixel p = row*W + col -> (u=col, v=H-1-row), int64."""
    ys, xs = np.meshgrid(np.arange(img_h), np.arange(img_w), indexing="ij")
    ys = (img_h - 1) - ys
    return np.stack([xs, ys], axis=-1).reshape(img_h * img_w, 2).astype(np.int64)


def ray_batch_partitions(n: int, num_batch: int) -> np.ndarray:
    """renderer/volume_renderer.py:229-235 -- torch.linspace(0, n, nb+1, dtype=long); last forced to n.

    torch computes integer linspace with a float64 step and truncates toward zero, using the
    symmetric form (start + step*i for the first half, end - step*(steps-1-i) for the rest).
    """
    steps = num_batch + 1
    step = (float(n) - 0.0) / float(steps - 1)
    half = steps // 2
    out = np.empty(steps, dtype=np.int64)
    for i in range(steps):
        v = 0.0 + step * i if i < half else float(n) - step * (steps - 1 - i)
        out[i] = int(v)
    out[-1] = n
    return out


# --------------------------------------------------------------------------------------
# K1 ray generation  (renderer/ray_samplers/sampler_base.py)
# --------------------------------------------------------------------------------------


def generate_rays(coords, intrinsic, extrinsic, t_near: float, img_h: int, img_w: int, project_to_ndc: bool):
    """sampler_base.py:134-197 (+ :70-113 directions, :115-132 origin, :199-257 NDC).

    coords (N,2) int64 screen coordinates.  Directions are NOT normalised (normalize=False,
    :159).  d = [x, y, -1] @ R^T (:164); o = 0 + c2w[:3, -1] (:165).
    """
    intrinsic = np.asarray(intrinsic, dtype=F32)
    extrinsic = np.asarray(extrinsic, dtype=F32)
    c = coords.astype(F32)
    x = ((c[:, 0] - intrinsic[0, 2]) / intrinsic[0, 0]).astype(F32)
    y = ((c[:, 1] - intrinsic[1, 2]) / intrinsic[1, 1]).astype(F32)
    d_cam = np.stack([x, y, -np.ones_like(x)], axis=-1).astype(F32)
    rot = extrinsic[:3, :3]
    ray_d = (d_cam @ rot.T).astype(F32)
    ray_o = (np.zeros_like(ray_d) + extrinsic[:3, -1][None, :]).astype(F32)
    if project_to_ndc:
        focal = float(intrinsic[0, 0])
        if float(intrinsic[0, 0]) != float(intrinsic[1, 1]):
            raise ValueError("Focal length used for computing NDC is ambiguous.")
        ray_o, ray_d = map_rays_to_ndc(focal, t_near, img_h, img_w, ray_o, ray_d)
    return ray_o, ray_d


def map_rays_to_ndc(focal: float, z_near: float, img_h: int, img_w: int, ray_o, ray_d):
    """sampler_base.py:199-257.  No shift of the origin to the near plane (reference quirk)."""
    if z_near < 0:
        raise ValueError(f"Expected a real number greater than or equal to 0. Got {z_near}.")
    # python-float scalars are applied to float32 tensors as float32 scalars
    sx = F32(-(2 * focal / img_w))
    sy = F32(-(2 * focal / img_h))
    two_n = F32(2 * z_near)
    with np.errstate(divide="ignore", invalid="ignore"):
        oxz = (ray_o[:, 0] / ray_o[:, 2]).astype(F32)
        oyz = (ray_o[:, 1] / ray_o[:, 2]).astype(F32)
        o_x = (sx * oxz).astype(F32)
        o_y = (sy * oyz).astype(F32)
        o_z = (F32(1) + (two_n / ray_o[:, 2]).astype(F32)).astype(F32)
        d_x = (sx * ((ray_d[:, 0] / ray_d[:, 2]).astype(F32) - oxz).astype(F32)).astype(F32)
        d_y = (sy * ((ray_d[:, 1] / ray_d[:, 2]).astype(F32) - oyz).astype(F32)).astype(F32)
        d_z = (-(two_n / ray_o[:, 2]).astype(F32)).astype(F32)
    return np.stack([o_x, o_y, o_z], -1).astype(F32), np.stack([d_x, d_y, d_z], -1).astype(F32)


# --------------------------------------------------------------------------------------
# K2 stratified coarse sampling  (renderer/ray_samplers/stratified_sampler.py)
# --------------------------------------------------------------------------------------


def create_t_bins(t_start: float, t_end: float, num_partitions: int):
    """stratified_sampler.py:130-164 -- torch.linspace(start, end, P+1)[:-1], step as python float.

    torch.linspace (float32, CPU) uses the symmetric scalar form start + step*i (i < steps/2)
    else end - step*(steps-1-i), step = (end-start)/(steps-1) in float32.  For the scene bounds
    the reference ships ((2,6) Blender, (0,1) NDC) every term is exact in float32.
    """
    steps = num_partitions + 1
    start, end = F32(t_start), F32(t_end)
    step = F32((end - start) / F32(steps - 1))
    half = steps // 2
    bins = np.empty(steps, dtype=F32)
    for i in range(steps):
        bins[i] = start + step * F32(i) if i < half else end - step * F32(steps - 1 - i)
    partition_size = (t_end - t_start) / num_partitions  # python float (float64)
    return bins[:-1].copy(), partition_size


def _deltas(t: np.ndarray) -> np.ndarray:
    """stratified_sampler.py:112-119 -- diff of [t, 1e8]; last interval is 1e8 - t_last."""
    far = np.full((t.shape[0], 1), 1e8, dtype=F32)
    return np.diff(np.concatenate([t, far], axis=-1), axis=-1).astype(F32)


def _points(ray_o, ray_d, t):
    """stratified_sampler.py:121-126 -- pts = o + t*d (mul then add); dirs replicated per sample."""
    n, s = t.shape
    dirs = np.repeat(ray_d[:, None, :], s, axis=1).astype(F32)
    pts = (ray_o[:, None, :] + (t[..., None] * dirs).astype(F32)).astype(F32)
    return pts, dirs


def sample_coarse_t(n: int, t_near: float, t_far: float, num_samples: int, u: np.ndarray) -> np.ndarray:
    """stratified_sampler.py:99-109 -- t = bins + step*u, u = rand_like (N,S)."""
    bins, step = create_t_bins(t_near, t_far, num_samples)
    return (bins[None, :] + (F32(step) * u.astype(F32)).astype(F32)).astype(F32)


def sample_along_rays_coarse(ray_o, ray_d, t_near, t_far, num_samples: int, u):
    """stratified_sampler.py:91-128 (coarse branch).  Returns pts (N,S,3), dirs (N,S,3), delta (N,S), t."""
    t = sample_coarse_t(ray_o.shape[0], t_near, t_far, num_samples, u)
    pts, dirs = _points(ray_o, ray_d, t)
    return pts, dirs, _deltas(t), t


# --------------------------------------------------------------------------------------
# K3 hierarchical sampling  (renderer/ray_samplers/utils.py, stratified_sampler.py)
# --------------------------------------------------------------------------------------


def torch_cpu_sum_lastdim(w: np.ndarray) -> np.ndarray:
    """Restates the accumulation order of torch.sum(w, dim=-1) on CPU float32 (utils.py:32).

    Pinned on the dev container (torch 2.11, AVX512 build): per 32-element block four
    interleaved 8-lane accumulators acc[a][l] += w[32c + 8a + l]; lanes combined as
    ((acc0+acc1)+acc2)+acc3; then the 8 lanes are added left to right.  Verified bit-exact for
    S in {8,16,32,64,96,128}; other S use the same model with a vector-8 then scalar tail
    and are NOT pinned.
    """
    w = np.ascontiguousarray(w, dtype=F32)
    n, s = w.shape
    acc = np.zeros((n, 4, 8), dtype=F32)
    nfull = s // 32
    for c in range(nfull):
        acc = (acc + w[:, 32 * c : 32 * (c + 1)].reshape(n, 4, 8)).astype(F32)
    tot = acc[:, 0]
    for a in range(1, 4):
        tot = (tot + acc[:, a]).astype(F32)
    off = nfull * 32
    while s - off >= 8:
        tot = (tot + w[:, off : off + 8]).astype(F32)
        off += 8
    r = tot[:, 0]
    for lane in range(1, 8):
        r = (r + tot[:, lane]).astype(F32)
    for i in range(off, s):
        r = (r + w[:, i]).astype(F32)
    return r


def pdf_to_cdf(weights: np.ndarray):
    """utils.py:31-40.  `weights` is modified IN PLACE (+= 1e-5), as in the reference.

    normaliser = torch.sum order (above); pdf = w / Z (IEEE float32 division); cdf = cumsum with a
    float64 running sum rounded to float32 per element (torch CPU cumsum accumulates in double),
    then shifted to an exclusive scan [0, c_0 .. c_{S-2}].
    """
    weights += F32(1e-5)
    z = torch_cpu_sum_lastdim(weights)
    pdf = (weights / z[:, None]).astype(F32)
    cdf_inc = np.cumsum(pdf.astype(np.float64), axis=-1).astype(F32)
    cdf = np.concatenate([np.zeros((cdf_inc.shape[0], 1), dtype=F32), cdf_inc[:, :-1]], axis=-1)
    return cdf


def sample_pdf(bins: np.ndarray, partition_size: float, weights: np.ndarray, u1: np.ndarray, u2: np.ndarray):
    """utils.py:8-58.  idx = searchsorted(cdf, u1, right=True) - 1 = #(cdf_j <= u) - 1 (integer,
    bit-exact gate); t = bins[idx] + step*u2 -- uniform inside the chosen bin, no interpolation.

    bins (N,S) float32; u1 = the torch.rand draw (N,F); u2 = the rand_like draw (N,F).
    Returns (t_fine (N,F) float32, idx (N,F) int64).
    """
    cdf = pdf_to_cdf(weights)
    u1 = u1.astype(F32)
    idx = (cdf[:, None, :] <= u1[:, :, None]).sum(axis=-1).astype(np.int64) - 1
    t_start = np.take_along_axis(bins, idx, axis=1)
    t = (t_start + (F32(partition_size) * u2.astype(F32)).astype(F32)).astype(F32)
    return t, idx


def sample_along_rays_fine(ray_o, ray_d, t_near, t_far, num_coarse: int, num_fine: int, weights, u0, u1, u2):
    """stratified_sampler.py:57-90 + :112-128.  A FRESH coarse draw (u0) is made (:77), fine samples
    come from sample_pdf (u1, u2), all 192 are sorted (:87-90).  `weights` is modified in place.

    Returns pts (N,S,3), dirs (N,S,3), delta (N,S), t (N,S) sorted, idx (N,F) int64.
    """
    bins, step = create_t_bins(t_near, t_far, num_coarse)
    bins2d = np.repeat(bins[None, :], ray_o.shape[0], axis=0)
    t_c = (bins2d + (F32(step) * u0.astype(F32)).astype(F32)).astype(F32)
    t_f, idx = sample_pdf(bins2d, step, weights, u1, u2)
    t = np.sort(np.concatenate([t_c, t_f], axis=-1), axis=-1).astype(F32)
    pts, dirs = _points(ray_o, ray_d, t)
    return pts, dirs, _deltas(t), t, idx


# --------------------------------------------------------------------------------------
# K4 positional encoding  (signal_encoder/positional_encoder.py)
# --------------------------------------------------------------------------------------


def positional_encode(x: np.ndarray, embed_level: int, include_input: bool = True) -> np.ndarray:
    """positional_encoder.py:49-104 -- [x | sin(2^0 x) | cos(2^0 x) | ... ], no pi; each function is
    applied to the whole C-vector, so the channel order is [xyz | sin f0 xyz | cos f0 xyz | ...]."""
    x = x.astype(F32)
    outs = [x] if include_input else []
    for lvl in range(embed_level):
        f = F32(2.0**lvl)
        fx = (f * x).astype(F32)
        outs.append(np.sin(fx).astype(F32))
        outs.append(np.cos(fx).astype(F32))
    return np.concatenate(outs, axis=-1).astype(F32)


# --------------------------------------------------------------------------------------
# K5/K6 NeRF MLP  (network/nerf.py)
# --------------------------------------------------------------------------------------

LAYER_NAMES = ["fc_in", "fc_1", "fc_2", "fc_3", "fc_4", "fc_5", "fc_6", "fc_7", "fc_8", "fc_9", "fc_out"]


def init_nerf_params(pos_dim: int = 63, view_dim: int = 27, feat: int = 256, seed: int = 0) -> dict:
    """Shapes of network/nerf.py:49-59 with nn.Linear-style U(-1/sqrt(in), 1/sqrt(in)) init (numpy RNG;
    NOT the torch init stream -- tests that need the reference's exact weights copy its state_dict)."""
    rng = np.random.default_rng(seed)
    dims = [(feat, pos_dim)] + [(feat, feat)] * 4 + [(feat, feat + pos_dim)] + [(feat, feat)] * 2
    dims += [(feat + 1, feat), (feat // 2, feat + view_dim), (3, feat // 2)]
    params = {}
    for name, (o, i) in zip(LAYER_NAMES, dims):
        b = 1.0 / np.sqrt(i)
        params[f"{name}.weight"] = rng.uniform(-b, b, size=(o, i)).astype(F32)
        params[f"{name}.bias"] = rng.uniform(-b, b, size=(o,)).astype(F32)
    return params


def _lin(x, params, name):
    return (x @ params[f"{name}.weight"].T + params[f"{name}.bias"]).astype(F32)


def nerf_forward(params: dict, pos: np.ndarray, view_dir: np.ndarray, return_cache: bool = False):
    """network/nerf.py:65-121.  pos (M,pos_dim) and view_dir (M,view_dim) are the ENCODED inputs.
    5 ReLU layers, cat [pos, h] -> fc_5 (:108), fc_6, fc_7, fc_8 without activation (:113),
    sigma = relu(out[:,0]) (:115), cat [out[:,1:], view_dir] -> fc_9 relu, fc_out sigmoid."""
    if pos.ndim != 2 or view_dir.ndim != 2:
        raise ValueError(f"Expected 2D tensors. Got {pos.ndim}, {view_dir.ndim}-D tensors.")
    if pos.shape[0] != view_dir.shape[0]:
        raise ValueError(f"The number of samples must match. Got {pos.shape[0]} and {view_dir.shape[0]}.")
    relu = lambda v: np.maximum(v, F32(0))
    acts = {}
    x = pos.astype(F32)
    acts["x_fc_in"] = x
    for i, name in enumerate(["fc_in", "fc_1", "fc_2", "fc_3", "fc_4"]):
        x = relu(_lin(x, params, name))
        if i < 4:
            acts[f"x_fc_{i + 1}"] = x
    x = np.concatenate([pos.astype(F32), x], axis=-1)
    acts["x_fc_5"] = x
    x = relu(_lin(x, params, "fc_5"))
    acts["x_fc_6"] = x
    x = relu(_lin(x, params, "fc_6"))
    acts["x_fc_7"] = x
    x = relu(_lin(x, params, "fc_7"))
    acts["x_fc_8"] = x
    out8 = _lin(x, params, "fc_8")
    acts["out8"] = out8
    sigma = relu(out8[:, 0])
    x = np.concatenate([out8[:, 1:], view_dir.astype(F32)], axis=-1)
    acts["x_fc_9"] = x
    x = relu(_lin(x, params, "fc_9"))
    acts["x_fc_out"] = x
    z = _lin(x, params, "fc_out")
    rgb = (F32(1) / (F32(1) + np.exp(-z))).astype(F32)
    acts["rgb"] = rgb
    if return_cache:
        return sigma, rgb, acts
    return sigma, rgb


def nerf_backward(params: dict, acts: dict, g_sigma: np.ndarray, g_rgb: np.ndarray) -> dict:
    """Gradient of nerf_forward w.r.t. the 22 parameter tensors (what autograd produces for
    network/nerf.py:102-119; inputs never require grad in the reference's callers)."""
    grads = {}

    def lin_bwd(name, x, g, need_dx=True):
        grads[f"{name}.weight"] = (g.T @ x).astype(F32)
        grads[f"{name}.bias"] = g.sum(axis=0).astype(F32)
        return (g @ params[f"{name}.weight"]).astype(F32) if need_dx else None

    rgb = acts["rgb"]
    g = (g_rgb * rgb * (F32(1) - rgb)).astype(F32)
    g = lin_bwd("fc_out", acts["x_fc_out"], g)
    g = (g * (acts["x_fc_out"] > 0)).astype(F32)
    g = lin_bwd("fc_9", acts["x_fc_9"], g)
    feat = params["fc_8.weight"].shape[0] - 1
    g8 = np.empty_like(acts["out8"])
    g8[:, 1:] = g[:, :feat]
    g8[:, 0] = g_sigma * (acts["out8"][:, 0] > 0)
    g = lin_bwd("fc_8", acts["x_fc_8"], g8)
    g = (g * (acts["x_fc_8"] > 0)).astype(F32)
    g = lin_bwd("fc_7", acts["x_fc_7"], g)
    g = (g * (acts["x_fc_7"] > 0)).astype(F32)
    g = lin_bwd("fc_6", acts["x_fc_6"], g)
    g = (g * (acts["x_fc_6"] > 0)).astype(F32)
    g = lin_bwd("fc_5", acts["x_fc_5"], g)
    pos_dim = acts["x_fc_in"].shape[1]
    g = g[:, pos_dim:]
    h4 = acts["x_fc_5"][:, pos_dim:]
    g = (g * (h4 > 0)).astype(F32)
    for i in (4, 3, 2, 1):
        g = lin_bwd(f"fc_{i}", acts[f"x_fc_{i}"], g)
        g = (g * (acts[f"x_fc_{i}"] > 0)).astype(F32)
    lin_bwd("fc_in", acts["x_fc_in"], g, need_dx=False)
    return grads


def query_points(params: dict, pts: np.ndarray, dirs: np.ndarray, l_pos: int = 10, l_dir: int = 4, return_cache=False):
    """scene/primitives/cube.py:39-76 -- flatten (n,S,3), encode both, network, reshape back."""
    if pts.shape != dirs.shape:
        raise ValueError(f"Expected tensors of same shape. Got {pts.shape} and {dirs.shape}, respectively.")
    n, s, _ = pts.shape
    pe = positional_encode(pts.reshape(n * s, -1), l_pos)
    de = positional_encode(dirs.reshape(n * s, -1), l_dir)
    if return_cache:
        sigma, rgb, acts = nerf_forward(params, pe, de, return_cache=True)
        return sigma.reshape(n, s), rgb.reshape(n, s, -1), acts
    sigma, rgb = nerf_forward(params, pe, de)
    return sigma.reshape(n, s), rgb.reshape(n, s, -1)


# --------------------------------------------------------------------------------------
# K7/K8 alpha compositing  (renderer/integrators/quadrature_integrator.py)
# --------------------------------------------------------------------------------------


def integrate_along_rays(sigma: np.ndarray, radiance: np.ndarray, delta: np.ndarray):
    """quadrature_integrator.py:14-67.  x = sigma*delta (:41); T_i = exp(-cumsum([0, x])[:-1]) (:44-52,
    torch CPU cumsum = float64 running sum rounded to float32 per element); alpha = 1 - exp(-x) (:55);
    w = T*alpha (:58); rgb = sum_i w_i c_i (:62-65).  A TRUE exclusive scan: never `inclusive - own`."""
    x = (sigma.astype(F32) * delta.astype(F32)).astype(F32)
    csum = np.cumsum(x.astype(np.float64), axis=-1).astype(F32)
    excl = np.concatenate([np.zeros((x.shape[0], 1), dtype=F32), csum[:, :-1]], axis=-1)
    with np.errstate(over="ignore", under="ignore"):
        trans = np.exp(-excl).astype(F32)
        alpha = (F32(1) - np.exp(-x).astype(F32)).astype(F32)
    w = (trans * alpha).astype(F32)
    rgb = (w[..., None] * radiance.astype(F32)).astype(F32).sum(axis=1).astype(F32)
    return rgb, w


def integrate_along_rays_backward(sigma, radiance, delta, g_rgb, g_w_ext=None):
    """Gradient of integrate_along_rays (autograd of quadrature_integrator.py:41-65), evaluated in
    float64 from the forward's definitions:
        g_c_i = w_i g_rgb;  g_w_i = g_rgb . c_i (+ external);  g_x_i = g_w_i (T_i - w_i) - sum_{k>i} g_w_k w_k;
        g_sigma_i = delta_i g_x_i.     (T_i - w_i = T_i exp(-x_i) = T_{i+1}.)
    The suffix sum is a true exclusive suffix scan."""
    s64 = sigma.astype(np.float64)
    d64 = delta.astype(np.float64)
    c64 = radiance.astype(np.float64)
    x = s64 * d64
    excl = np.concatenate([np.zeros((x.shape[0], 1)), np.cumsum(x, axis=-1)[:, :-1]], axis=-1)
    with np.errstate(over="ignore", under="ignore"):
        trans = np.exp(-excl)
        ex = np.exp(-x)
    w = trans * (1.0 - ex)
    g_rgb64 = g_rgb.astype(np.float64)
    g_c = w[..., None] * g_rgb64[:, None, :]
    g_w = (c64 * g_rgb64[:, None, :]).sum(-1)
    if g_w_ext is not None:
        g_w = g_w + g_w_ext.astype(np.float64)
    gw_w = g_w * w
    suffix_incl = np.cumsum(gw_w[:, ::-1], axis=-1)[:, ::-1]
    suffix_excl = np.concatenate([suffix_incl[:, 1:], np.zeros((x.shape[0], 1))], axis=-1)
    g_x = g_w * (trans * ex) - suffix_excl
    g_sigma = d64 * g_x
    return g_sigma.astype(F32), g_c.astype(F32)


# --------------------------------------------------------------------------------------
# L1 render pass + training-step gradients  (renderer/volume_renderer.py, runners/train.py)
# --------------------------------------------------------------------------------------


def render_pass(params, ray_o, ray_d, t_near, t_far, num_samples, uniforms, weights=None, num_ray_batch=None,
                return_cache=False):
    """renderer/volume_renderer.py:59-169 for given rays: sample (whole set), then chunked
    query + integrate (:229-254).  `num_samples` int -> coarse pass with uniforms=(u,);
    (Sc, Sf) -> fine pass with uniforms=(u0,u1,u2) and `weights` (modified in place)."""
    if isinstance(num_samples, (tuple, list)):
        sc, sf = num_samples
        pts, dirs, delta, t, idx = sample_along_rays_fine(ray_o, ray_d, t_near, t_far, sc, sf, weights, *uniforms)
    else:
        pts, dirs, delta, t = sample_along_rays_coarse(ray_o, ray_d, t_near, t_far, num_samples, uniforms[0])
        idx = None
    n = ray_o.shape[0]
    parts = ray_batch_partitions(n, 1 if num_ray_batch is None else num_ray_batch)
    rgbs, ws, caches = [], [], []
    for a, b in zip(parts[:-1], parts[1:]):
        if return_cache:
            sigma, rad, acts = query_points(params, pts[a:b], dirs[a:b], return_cache=True)
            caches.append((sigma, rad, acts))
        else:
            sigma, rad = query_points(params, pts[a:b], dirs[a:b])
        rgb, w = integrate_along_rays(sigma, rad, delta[a:b])
        rgbs.append(rgb)
        ws.append(w)
    out = dict(rgb=np.concatenate(rgbs, 0), weights=np.concatenate(ws, 0), t=t, delta=delta, idx=idx)
    if return_cache:
        out["cache"] = caches
    return out


def train_step_grads(params_c, params_f, ray_o, ray_d, t_near, t_far, sc, sf, target, u_c, u0, u1, u2):
    """One iteration of runners/train.py:130-218 up to loss.backward(): coarse render, MSE, fine render
    (fresh coarse draw + importance samples from the coarse weights, which carry no gradient --
    searchsorted/gather cut the graph, utils.py:47-55), MSE; returns losses and parameter grads of
    both networks.  MSELoss = mean over N*3 elements (runner_utils.py:731)."""
    n = ray_o.shape[0]
    co = render_pass(params_c, ray_o, ray_d, t_near, t_far, sc, (u_c,), return_cache=True)
    w_for_fine = co["weights"].copy()
    fi = render_pass(params_f, ray_o, ray_d, t_near, t_far, (sc, sf), (u0, u1, u2), weights=w_for_fine,
                     return_cache=True)
    out = {}
    for tag, res, params in (("coarse", co, params_c), ("fine", fi, params_f)):
        diff = (res["rgb"] - target).astype(F32)
        out[f"{tag}_loss"] = float(np.mean(diff.astype(np.float64) ** 2))
        g_rgb = (F32(2.0 / (n * 3)) * diff).astype(F32)
        sigma, rad, acts = res["cache"][0]
        g_sigma, g_c = integrate_along_rays_backward(sigma, rad, res["delta"], g_rgb)
        out[f"{tag}_grads"] = nerf_backward(params, acts, g_sigma.reshape(-1), g_c.reshape(-1, 3))
        out[f"{tag}_rgb"] = res["rgb"]
        out[f"{tag}_weights"] = res["weights"]
    out["idx"] = fi["idx"]
    return out


# --------------------------------------------------------------------------------------
# synthetic Blender-shaped cameras  (utils/data/load_blender.py)
# --------------------------------------------------------------------------------------


def pose_spherical(theta_deg: float, phi_deg: float, radius: float) -> np.ndarray:
    """utils/data/load_blender.py:78-109 -- c2w = flip @ (rot_y(theta) @ (rot_x(phi) @ trans_z(radius))),
    every factor a float32 matrix and every product a float32 matmul, as in the reference."""
    t = np.eye(4, dtype=F32)
    t[2, 3] = radius
    phi = phi_deg / 180.0 * np.pi
    rp = np.array([[1, 0, 0, 0], [0, np.cos(phi), -np.sin(phi), 0], [0, np.sin(phi), np.cos(phi), 0], [0, 0, 0, 1.0]],
                  dtype=F32)
    th = theta_deg / 180.0 * np.pi
    rt = np.array([[np.cos(th), 0, -np.sin(th), 0], [0, 1, 0, 0], [np.sin(th), 0, np.cos(th), 0], [0, 0, 0, 1.0]],
                  dtype=F32)
    flip = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1.0]], dtype=F32)
    return (flip @ (rt @ (rp @ t).astype(F32)).astype(F32)).astype(F32)


def blender_focal(img_w: int, camera_angle_x: float = 0.6911112070083618) -> float:
    """utils/data/load_blender.py:170-171 -- focal = 0.5 W / tan(0.5 camera_angle_x)."""
    return float(0.5 * img_w / np.tan(0.5 * camera_angle_x))


# ------------------------------------------------------------------------------------------------
# callers of the path (section 8f): the pieces of the training loop that are arithmetic
# ------------------------------------------------------------------------------------------------
def center_crop_pixel_indices(img_h: int, img_w: int) -> np.ndarray:
    """runners/train.py:151-167 -- flat pixel ids (row * W + col) of the central block sampled while epoch < 10:
    the cartesian product of arange(ci - ci//2, ci + ci//2) x arange(cj - cj//2, cj + cj//2), rows outermost."""
    ci, cj = (img_h - 1) // 2, (img_w - 1) // 2
    out = []
    for i in range(ci - ci // 2, ci + ci // 2):
        for j in range(cj - cj // 2, cj + cj // 2):
            out.append(i * img_w + j)
    return np.asarray(out, dtype=np.int64)


def exp_lr_gamma(init_lr: float, end_lr: float, num_iter: int) -> float:
    """runners/runner_utils.py:704-708 -- per-iteration decay of ExponentialLR."""
    return pow(end_lr / init_lr, 1 / num_iter)


def adam_step(p, g, m, v, step: int, lr: float, eps: float = 1e-8, b1: float = 0.9, b2: float = 0.999):
    """torch.optim.Adam (runner_utils.py:691-695: lr, eps given; betas, weight_decay = defaults), single tensor,
    fp32 state like torch: returns the new (p, m, v) after update number `step` (1-based)."""
    F = np.float32
    m = (F(b1) * m + F(1 - b1) * g).astype(F)
    v = (F(b2) * v + F(1 - b2) * g * g).astype(F)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = (np.sqrt(v) / F(np.sqrt(bc2)) + F(eps)).astype(F)
    p = (p - F(lr / bc1) * (m / denom)).astype(F)
    return p, m, v
