"""
Generates the golden vectors under tests/golden/ by running the UNMODIFIED reference modules
(imported from /root/reference, CPU) on small seeded inputs.  Run in the dev container only:

    python tests/golden/make_golden.py

The reference has no tests/fixtures of its own, so these frozen outputs are what pins the
oracle (oracle/nerf_oracle.py) and, through it, the CUDA path.  Network parameters are NOT
stored: they are regenerated from a numpy seed with oracle.init_nerf_params and loaded into the
reference's NeRF module here, so fixtures stay small.  Uniform draws are replayed into the
reference by temporarily replacing torch.rand / torch.rand_like with a queue.
"""

import os
import sys
from contextlib import contextmanager

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

from torch_nerf.src.network.nerf import NeRF  # noqa: E402
from torch_nerf.src.renderer.cameras import PerspectiveCamera  # noqa: E402
from torch_nerf.src.renderer.integrators.quadrature_integrator import QuadratureIntegrator  # noqa: E402
from torch_nerf.src.renderer.ray_samplers.stratified_sampler import StratifiedSampler  # noqa: E402
from torch_nerf.src.renderer.ray_samplers.utils import sample_pdf  # noqa: E402
from torch_nerf.src.renderer.ray_samplers.sampler_base import RayBundle  # noqa: E402
from torch_nerf.src.renderer.volume_renderer import VolumeRenderer  # noqa: E402
from torch_nerf.src.scene.primitives.cube import PrimitiveCube  # noqa: E402
from torch_nerf.src.signal_encoder.positional_encoder import PositionalEncoder  # noqa: E402

from oracle import nerf_oracle as orc  # noqa: E402

torch.set_num_threads(1)


@contextmanager
def replay_uniforms(queue):
    """Feeds the given arrays, in order, to torch.rand / torch.rand_like calls."""
    q = [torch.from_numpy(np.ascontiguousarray(a)) for a in queue]
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


def load_params(net: NeRF, params: dict):
    net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in params.items()})
    return net


def make_scene(seed: int):
    params = orc.init_nerf_params(seed=seed)
    net = load_params(NeRF(63, 27), params)
    enc = {"coord_enc": PositionalEncoder(3, 10, True), "dir_enc": PositionalEncoder(3, 4, True)}
    return PrimitiveCube(net, enc), net, params


def grad_digest(named_grads: dict, rng_seed: int = 123):
    """Per-tensor (sum, abs-sum, 64 sampled entries at fixed positions) so fixtures stay small."""
    out = {}
    rng = np.random.default_rng(rng_seed)
    for k in sorted(named_grads):
        g = np.asarray(named_grads[k], dtype=np.float32).reshape(-1)
        pos = rng.integers(0, g.size, size=min(64, g.size))
        out[f"{k}/sum"] = np.float64(g.astype(np.float64).sum())
        out[f"{k}/abssum"] = np.float64(np.abs(g.astype(np.float64)).sum())
        out[f"{k}/pos"] = pos.astype(np.int64)
        out[f"{k}/val"] = g[pos]
    return out


def camera(h, w, focal, c2w, near, far):
    return PerspectiveCamera({"f_x": focal, "f_y": focal, "img_width": w, "img_height": h}, torch.from_numpy(c2w), near, far)


def save(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrs)
    print(f"wrote {name}: {os.path.getsize(path) / 1024:.1f} KiB")


def gen_raygen():
    rng = np.random.default_rng(1)
    sampler = StratifiedSampler()
    # Blender-shaped 800x800 camera, random pixels
    h = w = 800
    focal = orc.blender_focal(w)
    c2w = orc.pose_spherical(37.0, -30.0, 4.0)
    cam = camera(h, w, focal, c2w, 2.0, 6.0)
    ren = VolumeRenderer(QuadratureIntegrator(), sampler, cam)
    pix = rng.choice(h * w, size=64, replace=False).astype(np.int64)
    coords = ren.screen_coords.clone()[torch.from_numpy(pix), :]
    b = sampler.generate_rays(coords.clone(), cam, project_to_ndc=False)
    # LLFF-shaped camera, NDC, near in {0.0 (shipped, degenerate), 1.0}
    h2, w2, f2 = 756, 1008, 815.13
    c2w2 = np.eye(4, dtype=np.float32)
    a = 0.05
    c2w2[:3, :3] = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]], dtype=np.float32)
    c2w2[:3, 3] = np.array([0.13, -0.07, 0.21], dtype=np.float32)
    pix2 = rng.choice(h2 * w2, size=64, replace=False).astype(np.int64)
    outs = {}
    for tag, near in (("ndc0", 0.0), ("ndc1", 1.0)):
        cam2 = camera(h2, w2, f2, c2w2[:3, :], near, 1.0)
        ren2 = VolumeRenderer(QuadratureIntegrator(), sampler, cam2)
        coords2 = ren2.screen_coords.clone()[torch.from_numpy(pix2), :]
        b2 = sampler.generate_rays(coords2.clone(), cam2, project_to_ndc=True)
        outs[f"{tag}_o"] = b2.ray_origin.numpy()
        outs[f"{tag}_d"] = b2.ray_dir.numpy()
        outs[f"{tag}_coords"] = coords2.numpy()
    save("raygen.npz", h=h, w=w, focal=np.float64(focal), c2w=c2w, pix=pix, coords=coords.numpy(),
         ray_o=b.ray_origin.numpy(), ray_d=b.ray_dir.numpy(),
         h2=h2, w2=w2, focal2=np.float64(f2), c2w2=c2w2[:3, :], pix2=pix2, **outs)


def gen_sampling():
    rng = np.random.default_rng(2)
    sampler = StratifiedSampler()
    n, sc, sf = 48, 64, 128
    ray_o = rng.normal(size=(n, 3)).astype(np.float32)
    ray_d = rng.normal(size=(n, 3)).astype(np.float32)
    u = rng.random((n, sc), dtype=np.float32)
    res = {}
    for tag, (near, far) in (("b", (2.0, 6.0)), ("n", (0.0, 1.0))):
        bundle = RayBundle(torch.from_numpy(ray_o), torch.from_numpy(ray_d), near, far, False)
        with replay_uniforms([u]):
            pts, dirs, delta = sampler.sample_along_rays(bundle, sc, device="cpu")
        res[f"{tag}_pts"], res[f"{tag}_dirs"], res[f"{tag}_delta"] = pts.numpy(), dirs.numpy(), delta.numpy()
    save("coarse.npz", ray_o=ray_o, ray_d=ray_d, u=u, **res)

    # hierarchical: realistic peaked weights (some rays nearly empty, some with one dominant bin)
    w = (rng.random((n, sc), dtype=np.float32) ** 6).astype(np.float32)
    w[:8] *= 1e-4
    w[8:16, 20] = 0.9
    w[16:20] = 0.0
    u0 = rng.random((n, sc), dtype=np.float32)
    u1 = rng.random((n, sf), dtype=np.float32)
    u2 = rng.random((n, sf), dtype=np.float32)
    u1[0, :4] = 0.0  # edge: u == cdf[0]
    res = {}
    for tag, (near, far) in (("b", (2.0, 6.0)), ("n", (0.0, 1.0))):
        bundle = RayBundle(torch.from_numpy(ray_o), torch.from_numpy(ray_d), near, far, False)
        w_in = torch.from_numpy(w.copy())
        with replay_uniforms([u0, u1, u2]):
            pts, dirs, delta = sampler.sample_along_rays(bundle, (sc, sf), device="cpu", weights=w_in)
        res[f"{tag}_pts"], res[f"{tag}_delta"] = pts.numpy(), delta.numpy()
        res[f"{tag}_w_after"] = w_in.numpy()  # in-place += 1e-5 side effect
        # bin indices: re-run utils.py:31-54 verbatim
        wt = torch.from_numpy(w.copy())
        wt += 1e-5
        pdf = wt / torch.sum(wt, dim=-1, keepdim=True)
        cdf = torch.cumsum(pdf, dim=-1)
        cdf = torch.cat([torch.zeros((cdf.shape[0], 1)), cdf[..., :-1]], dim=-1)
        idx = torch.searchsorted(cdf, torch.from_numpy(u1).contiguous(), right=True) - 1
        res[f"{tag}_idx"] = idx.numpy()
        # and sample_pdf itself
        t_bins = torch.linspace(near, far, sc + 1)[:-1].unsqueeze(0).repeat(n, 1)
        with replay_uniforms([u1, u2]):
            t_f = sample_pdf(t_bins, (far - near) / sc, torch.from_numpy(w.copy()), sf)
        res[f"{tag}_t_fine"] = t_f.numpy()
    save("fine.npz", ray_o=ray_o, ray_d=ray_d, weights=w, u0=u0, u1=u1, u2=u2, **res)

    # larger bit-exactness check of the bin indices only (4096 rays), stored as a checksum
    n2 = 4096
    rng_big = np.random.default_rng(202)  # the test regenerates w2/u1b from this seed
    w2 = (rng_big.random((n2, sc), dtype=np.float32) ** 4 * rng_big.random((n2, 1), dtype=np.float32)).astype(np.float32)
    u1b = rng_big.random((n2, sf), dtype=np.float32)
    wt = torch.from_numpy(w2.copy())
    wt += 1e-5
    pdf = wt / torch.sum(wt, dim=-1, keepdim=True)
    cdf = torch.cumsum(pdf, dim=-1)
    cdf = torch.cat([torch.zeros((n2, 1)), cdf[..., :-1]], dim=-1)
    idx = (torch.searchsorted(cdf, torch.from_numpy(u1b).contiguous(), right=True) - 1).numpy()
    save("fine_idx_big.npz", seed=202, idx=idx.astype(np.uint8), cdf_last=cdf[:, -1].numpy())


def gen_posenc():
    rng = np.random.default_rng(3)
    x = (rng.normal(size=(40, 3)) * 2.5).astype(np.float32)
    x[0] = 0.0
    x[1] = [6.0, -6.0, 5.9]
    out10 = PositionalEncoder(3, 10, True).encode(torch.from_numpy(x)).numpy()
    out4 = PositionalEncoder(3, 4, True).encode(torch.from_numpy(x)).numpy()
    out4n = PositionalEncoder(3, 4, False).encode(torch.from_numpy(x)).numpy()
    save("posenc.npz", x=x, out10=out10, out4=out4, out4_noinput=out4n)


def gen_mlp():
    rng = np.random.default_rng(4)
    scene, net, params = make_scene(seed=11)
    m = 96
    pts = (rng.normal(size=(m, 3)) * 1.5).astype(np.float32)
    dirs = rng.normal(size=(m, 3)).astype(np.float32)
    pe = PositionalEncoder(3, 10, True).encode(torch.from_numpy(pts))
    de = PositionalEncoder(3, 4, True).encode(torch.from_numpy(dirs))
    sigma, rgb = net(pe, de)
    g_sigma = rng.normal(size=(m,)).astype(np.float32)
    g_rgb = rng.normal(size=(m, 3)).astype(np.float32)
    loss = (sigma * torch.from_numpy(g_sigma)).sum() + (rgb * torch.from_numpy(g_rgb)).sum()
    loss.backward()
    grads = {k: p.grad.numpy() for k, p in net.named_parameters()}
    save("mlp.npz", seed=11, pts=pts, dirs=dirs, pe=pe.numpy(), de=de.numpy(), sigma=sigma.detach().numpy(),
         rgb=rgb.detach().numpy(), g_sigma=g_sigma, g_rgb=g_rgb, **grad_digest(grads))


def gen_composite():
    rng = np.random.default_rng(5)
    n, s = 24, 192
    sigma = np.maximum(rng.normal(size=(n, s)) * 3.0, 0).astype(np.float32)
    sigma[0] = 0.0
    sigma[1] = 50.0
    rad = rng.random((n, s, 3), dtype=np.float32)
    t = np.sort(2.0 + 4.0 * rng.random((n, s)), axis=-1).astype(np.float32)
    delta = np.diff(np.concatenate([t, np.full((n, 1), 1e8, np.float32)], -1), axis=-1).astype(np.float32)
    sg = torch.from_numpy(sigma).requires_grad_(True)
    rd = torch.from_numpy(rad).requires_grad_(True)
    rgb, w = QuadratureIntegrator().integrate_along_rays(sg, rd, torch.from_numpy(delta))
    g_rgb = rng.normal(size=(n, 3)).astype(np.float32)
    g_w = rng.normal(size=(n, s)).astype(np.float32)
    (rgb * torch.from_numpy(g_rgb)).sum().backward(retain_graph=True)
    gs0, gr0 = sg.grad.numpy().copy(), rd.grad.numpy().copy()
    sg.grad = None
    rd.grad = None
    ((rgb * torch.from_numpy(g_rgb)).sum() + (w * torch.from_numpy(g_w)).sum()).backward()
    save("composite.npz", sigma=sigma, radiance=rad, delta=delta, rgb=rgb.detach().numpy(), w=w.detach().numpy(),
         g_rgb=g_rgb, g_w=g_w, g_sigma=gs0, g_radiance=gr0, g_sigma_w=sg.grad.numpy(), g_radiance_w=rd.grad.numpy())


def gen_render():
    """Config C1 in miniature: whole-frame coarse + fine render of a 12x10 Blender-shaped view."""
    rng = np.random.default_rng(6)
    h, w = 10, 12
    focal = orc.blender_focal(w)
    c2w = orc.pose_spherical(-60.0, -30.0, 4.0)
    cam = camera(h, w, focal, c2w, 2.0, 6.0)
    scene_c, _, _ = make_scene(seed=21)
    scene_f, _, _ = make_scene(seed=22)
    ren = VolumeRenderer(QuadratureIntegrator(), StratifiedSampler(), cam)
    n, sc, sf = h * w, 64, 128
    u_c = rng.random((n, sc), dtype=np.float32)
    u0 = rng.random((n, sc), dtype=np.float32)
    u1 = rng.random((n, sf), dtype=np.float32)
    u2 = rng.random((n, sf), dtype=np.float32)
    with torch.no_grad():
        with replay_uniforms([u_c]):
            rgb_c, pix, w_c = ren.render_scene(scene_c, n, sc, False, "cpu", num_ray_batch=2)
        w_c_before = w_c.numpy().copy()
        with replay_uniforms([u0, u1, u2]):
            rgb_f, _, w_f = ren.render_scene(scene_f, n, (sc, sf), False, "cpu", pixel_indices=pix, weights=w_c,
                                             num_ray_batch=2)
    save("render.npz", h=h, w=w, focal=np.float64(focal), c2w=c2w, seed_c=21, seed_f=22, u_c=u_c, u0=u0, u1=u1, u2=u2,
         pix=pix.numpy(), rgb_c=rgb_c.numpy(), w_c=w_c_before, rgb_f=rgb_f.numpy(), w_f=w_f.numpy())


def gen_train_step():
    """Config C2 in miniature: one iteration of train.py:130-218 (up to backward) on 64 rays of an
    800x800 camera."""
    rng = np.random.default_rng(7)
    h = w = 800
    focal = orc.blender_focal(w)
    c2w = orc.pose_spherical(110.0, -25.0, 4.0)
    cam = camera(h, w, focal, c2w, 2.0, 6.0)
    scene_c, net_c, _ = make_scene(seed=31)
    scene_f, net_f, _ = make_scene(seed=32)
    ren = VolumeRenderer(QuadratureIntegrator(), StratifiedSampler(), cam)
    n, sc, sf = 64, 64, 128
    pix = torch.from_numpy(rng.choice(h * w, size=n, replace=False).astype(np.int64))
    target = rng.random((n, 3), dtype=np.float32)
    u_c = rng.random((n, sc), dtype=np.float32)
    u0 = rng.random((n, sc), dtype=np.float32)
    u1 = rng.random((n, sf), dtype=np.float32)
    u2 = rng.random((n, sf), dtype=np.float32)
    loss_fn = torch.nn.MSELoss()
    with replay_uniforms([u_c]):
        pred_c, idx_c, w_c = ren.render_scene(scene_c, n, sc, False, "cpu", pixel_indices=pix)
    loss_c = loss_fn(torch.from_numpy(target), pred_c)
    with replay_uniforms([u0, u1, u2]):
        pred_f, _, w_f = ren.render_scene(scene_f, n, (sc, sf), False, "cpu", pixel_indices=idx_c, weights=w_c)
    loss_f = loss_fn(torch.from_numpy(target), pred_f)
    (loss_c + loss_f).backward()
    gc = {k: p.grad.numpy() for k, p in net_c.named_parameters()}
    gf = {k: p.grad.numpy() for k, p in net_f.named_parameters()}
    dig = {f"c/{k}": v for k, v in grad_digest(gc).items()}
    dig.update({f"f/{k}": v for k, v in grad_digest(gf).items()})
    save("train_step.npz", h=h, w=w, focal=np.float64(focal), c2w=c2w, seed_c=31, seed_f=32, pix=pix.numpy(),
         target=target, u_c=u_c, u0=u0, u1=u1, u2=u2, loss_c=np.float64(loss_c.item()), loss_f=np.float64(loss_f.item()),
         rgb_c=pred_c.detach().numpy(), rgb_f=pred_f.detach().numpy(), w_f=w_f.detach().numpy(), **dig)


if __name__ == "__main__":
    gen_raygen()
    gen_sampling()
    gen_posenc()
    gen_mlp()
    gen_composite()
    gen_render()
    gen_train_step()
