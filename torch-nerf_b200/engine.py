"""Fused hot-path engine: the coarse -> fine render pass and the training step (forward + backward) driven as one
stream of libnerf_b200 kernels with no host synchronisation and no (N,S,3) materialisation.

This is the fast caller of the same C-ABI the drop-in classes use.  It restates what the reference's callers do
around `VolumeRenderer.render_scene`:
  * training iteration  -- runners/train.py:130-218 (coarse render, MSE, fine render from the coarse weights, MSE,
                           backward; optimizer step stays with the caller)
  * full-frame render   -- runners/render.py:58-107 (coarse + fine over all H*W pixels, clamp to [0,1])
Uniform draws follow the reference's order and shapes (1 draw for the coarse pass, 3 for the fine pass).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _lib
from .cameras import PerspectiveCamera
from .network import NeRF


class FlatParams:
    """Re-homes the parameters of the coarse and fine networks into ONE flat fp32 buffer (and one flat gradient
    buffer), so the data-parallel exchange is a single all-reduce over 2 x 595 844 floats and the kernels write
    gradients straight into it."""

    def __init__(self, nets: Sequence[NeRF], grad_buffer: Optional[torch.Tensor] = None):
        self.nets = list(nets)
        params = [p for net in self.nets for p in net.ordered_parameters()]
        dev = params[0].device
        total = sum(p.numel() for p in params)
        self.flat = torch.empty(total, device=dev, dtype=torch.float32)
        if grad_buffer is None:
            self.grad = torch.zeros(total, device=dev, dtype=torch.float32)
        else:  # e.g. parallel.PeerExchange.grad: symmetric memory the other ranks read over NVLink
            if grad_buffer.shape != (total,) or grad_buffer.dtype != torch.float32 or grad_buffer.device != dev:
                raise ValueError(f"grad_buffer must be a float32 tensor of {total} elements on {dev}")
            self.grad = grad_buffer
            self.grad.zero_()
        off = 0
        self.grad_views = []
        for p in params:
            n = p.numel()
            self.flat[off:off + n].copy_(p.detach().reshape(-1))
            p.data = self.flat[off:off + n].view_as(p)
            g = self.grad[off:off + n].view_as(p)
            p.grad = g
            self.grad_views.append(g)
            off += n
        self.per_net = len(params) // len(self.nets)
        # one Parameter over the whole buffer: an optimizer stepping it updates every network parameter (they alias it)
        self.param = torch.nn.Parameter(self.flat, requires_grad=True)
        self.param.grad = self.grad

    def grads_of(self, i: int):
        return self.grad_views[i * self.per_net:(i + 1) * self.per_net]


class _StepGraph:
    """Static inputs and the captured graph of one training-iteration shape (see HotPathEngine.train_pixels_graph)."""

    def __init__(self, n: int, near: float, far: float, device):
        import ctypes

        self.n, self.near, self.far = n, near, far
        self.pix = torch.empty((n,), device=device, dtype=torch.int64)
        self.tgt = torch.empty((n, 3), device=device, dtype=torch.float32)
        self.cam = torch.zeros((ctypes.sizeof(_lib.CameraStruct),), device=device, dtype=torch.uint8)
        self.losses = torch.zeros((2,), device=device, dtype=torch.float32)
        self.graph = None
        self.launches = 0
        self.keepalive = []  # every workspace the captured kernels address (the engine may outgrow and replace them)


class HotPathEngine:
    def __init__(self, coarse: NeRF, fine: NeRF, num_coarse: int = 64, num_fine: int = 128, precision: str = "bf16"):
        self.lib = _lib.load()
        self.nets = (coarse, fine)
        self.sc, self.sf = int(num_coarse), int(num_fine)
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        if precision == "bf16" and not (coarse.supports_bf16() and fine.supports_bf16()):
            raise ValueError("the bf16 tensor-core chain is built for NeRF(63, 27, 256)")
        self.precision = precision
        self.device = next(coarse.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("HotPathEngine needs CUDA parameters (no CPU fallback)")
        self.flat: Optional[FlatParams] = None
        self._buf: Dict[Tuple, torch.Tensor] = {}
        self._graphs: Dict[Tuple, _StepGraph] = {}
        self.launches = 0  # kernels of this library enqueued so far (host-side count)
        self._prologue_done = False  # inside train_rays: both bf16 weight images are fresh, gradients and losses zeroed

    # ------------------------------------------------------------------------------------------ buffers
    def _get(self, name: str, shape, dtype=torch.float32) -> torch.Tensor:
        """Workspace `name` as a tensor of `shape`: one grow-only allocation per name (a smaller request is a view of
        it), so a warm-up batch, a last shard or another frame size do not each pin their own multi-GB set."""
        shape = tuple(int(x) for x in shape)
        numel = 1
        for x in shape:
            numel *= x
        key = (name, dtype)
        t = self._buf.get(key)
        if t is None or t.numel() < numel:
            t = torch.empty((numel,), device=self.device, dtype=dtype)
            self._buf[key] = t
        return t[:numel].view(shape)

    def release_workspaces(self) -> None:
        """Drops every workspace and captured iteration (they are re-created on demand)."""
        self._buf.clear()
        self._graphs.clear()

    def _call(self, fn_name: str, *args, launches: int = 1):
        _lib.check(getattr(self.lib, fn_name)(*args), fn_name)
        self.launches += launches

    def enable_flat_params(self, grad_buffer: Optional[torch.Tensor] = None) -> FlatParams:
        if self.flat is None:
            self.flat = FlatParams(self.nets, grad_buffer)
        elif grad_buffer is not None and grad_buffer.data_ptr() != self.flat.grad.data_ptr():
            raise RuntimeError("flat parameters already exist with another gradient buffer")
        return self.flat

    # ------------------------------------------------------------------------------------------ pieces
    def _uniforms(self, n: int, uniforms, fine: bool):
        if uniforms is not None:
            return [u.to(self.device, torch.float32).contiguous() for u in uniforms]
        if fine:
            return [torch.rand((n, self.sc), device=self.device), torch.rand((n, self.sf), device=self.device),
                    torch.rand((n, self.sf), device=self.device)]
        return [torch.rand((n, self.sc), device=self.device)]

    def _sample(self, tag, ray_o, ray_d, n, near, far, weights, u, need_points: bool):
        P, st = _lib.ptr, _lib.stream()
        s = self.sc + self.sf if weights is not None else self.sc
        t = self._get(tag + "t", (n, s))
        delta = self._get(tag + "delta", (n, s))
        pts = self._get(tag + "pts", (n, s, 3)) if need_points else None
        dirs = self._get(tag + "dirs", (n, s, 3)) if need_points else None
        if weights is None:
            self._call("nerf_sample_coarse", P(ray_o), P(ray_d), n, self.sc, near, far, P(u[0]), P(t), P(pts), P(dirs),
                       P(delta), st)
        else:
            self._call("nerf_sample_fine", P(ray_o), P(ray_d), n, self.sc, self.sf, near, far, P(weights), P(u[0]), P(u[1]),
                       P(u[2]), None, P(t), P(pts), P(dirs), P(delta), st)
        return t, delta, pts, dirs, s

    def _pass(self, which: int, tag: str, ray_o, ray_d, n, near, far, weights, u, train: bool, want_depth: bool = False):
        """One render pass (coarse: which=0, fine: which=1): sample -> query -> composite.  `want_depth` also produces
        depth = sum_i w_i t_i and opacity = sum_i w_i (north-star outputs without a reference counterpart)."""
        P, st = _lib.ptr, _lib.stream()
        net = self.nets[which]
        fp32 = self.precision == "fp32"
        t, delta, pts, dirs, s = self._sample(tag, ray_o, ray_d, n, near, far, weights, u, need_points=fp32)
        m = n * s
        sigma = self._get(tag + "sigma", (n, s))
        rad = self._get(tag + "rad", (n, s, 3))
        ctx = {"t": t, "delta": delta, "sigma": sigma, "rad": rad, "s": s, "m": m}
        if fp32:
            pe = self._get(tag + "pe", (m, net.pos_dim))
            de = self._get(tag + "de", (m, net.view_dir_dim))
            lp = (net.pos_dim - 3) // 6
            lv = (net.view_dir_dim - 3) // 6
            self._call("nerf_posenc", P(pts), m, 3, lp, 1, P(pe), net.pos_dim, st)
            self._call("nerf_posenc", P(dirs), m, 3, lv, 1, P(de), net.view_dir_dim, st)
            cache = self._get(tag + "cache", (self.lib.nerf_mlp_f32_cache_floats(net._dims, m),))
            params = _lib.pointer_array([p.detach() for p in net.ordered_parameters()])
            self._call("nerf_mlp_f32_forward", net._dims, params, P(pe), P(de), m, P(sigma), P(rad), P(cache), st, launches=16)
            ctx["cache"] = cache
        else:
            # with flat parameters an optimizer may have stepped the shared buffer without touching the per-tensor
            # version counters, so the bf16 image is rebuilt on every pass (2.3 MB, ~10 us)
            # (a training iteration packs both networks in its one-launch prologue: nothing to do here then)
            packed = net._packed if self._prologue_done else net.packed_weights(force=train or self.flat is not None)
            cache = None
            if train:
                cache = self._get(tag + "cache16", (self.lib.nerf_mlp_bf16_cache_bytes(m),), torch.uint8)
                ctx["cache"] = cache
            self._call("nerf_mlp_bf16_forward", P(packed, torch.uint8), None, None, P(ray_o), P(ray_d), P(t), s, m, P(sigma),
                       P(rad), P(cache, torch.uint8) if cache is not None else None, st)
        rgb = self._get(tag + "rgb", (n, 3))
        w = self._get(tag + "w", (n, s))
        depth = self._get(tag + "depth", (n,)) if want_depth else None
        opacity = self._get(tag + "opacity", (n,)) if want_depth else None
        self._call("nerf_composite_fwd", P(sigma), P(rad), P(delta), P(t) if want_depth else None, n, s, P(rgb), P(w), P(depth),
                   P(opacity), st)
        ctx["rgb"], ctx["w"], ctx["depth"], ctx["opacity"] = rgb, w, depth, opacity
        return ctx

    def _backward(self, which: int, ctx, n: int, g_rgb, grads, target=None, loss=None):
        """Backward of one pass.  With `target` the MSE head is folded into the compositing backward (g_rgb is never
        materialised, `loss` += the pass's MSE); otherwise `g_rgb` is the upstream gradient."""
        P, st = _lib.ptr, _lib.stream()
        net = self.nets[which]
        s, m = ctx["s"], ctx["m"]
        tag = "bc" if which == 0 else "bf"
        g_sigma = self._get(tag + "gs", (n, s))
        g_rad = self._get(tag + "gr", (n, s, 3))
        if target is not None and s <= 256:
            self._call("nerf_composite_bwd_mse", P(ctx["sigma"]), P(ctx["rad"]), P(ctx["delta"]), P(ctx["rgb"]), P(target), n, s,
                       P(g_sigma), P(g_rad), P(loss), st)
        else:
            if target is not None:
                g_rgb = self._get(tag + "g_rgb", (n, 3))
                self._call("nerf_mse_loss", P(ctx["rgb"]), P(target), n, P(g_rgb), P(loss), st)
            self._call("nerf_composite_bwd", P(ctx["sigma"]), P(ctx["rad"]), P(ctx["delta"]), P(g_rgb), None, n, s, P(g_sigma),
                       P(g_rad), st)
        garr = _lib.pointer_array(grads)
        if self.precision == "fp32":
            scratch = self._get("scratch32", (self.lib.nerf_mlp_f32_bwd_scratch_floats(net._dims, n * (self.sc + self.sf)),))
            params = _lib.pointer_array([p.detach() for p in net.ordered_parameters()])
            self._call("nerf_mlp_f32_backward", net._dims, params, P(ctx["cache"]), P(ctx["rad"]), m, P(g_sigma), P(g_rad),
                       garr, P(scratch), st, launches=60)
        else:
            scratch = self._get("scratch16", (self.lib.nerf_mlp_bf16_bwd_scratch_bytes(n * (self.sc + self.sf)),), torch.uint8)
            if self._prologue_done:  # gradients were zeroed by the prologue: chain + weight gradients only
                self._call("nerf_mlp_bf16_backward_part", P(net._packed, torch.uint8), P(ctx["cache"], torch.uint8), P(ctx["rad"]),
                           m, P(g_sigma), P(g_rad), garr, P(scratch, torch.uint8), 6, 0, (m + 127) // 128, 0, st, launches=2)
            else:
                self._call("nerf_mlp_bf16_backward", P(net.packed_weights(), torch.uint8), P(ctx["cache"], torch.uint8),
                           P(ctx["rad"]), m, P(g_sigma), P(g_rad), garr, P(scratch, torch.uint8), st, launches=3)

    # ------------------------------------------------------------------------------------------ public
    def rays_from_pixels(self, camera: PerspectiveCamera, project_to_ndc: bool, pixel_indices: Optional[torch.Tensor],
                         first_pixel: int = 0, count: int = 0):
        P, st = _lib.ptr, _lib.stream()
        n = int(pixel_indices.shape[0]) if pixel_indices is not None else int(count)
        ray_o = self._get("ray_o", (n, 3))
        ray_d = self._get("ray_d", (n, 3))
        cam = camera.pack(project_to_ndc)
        self._call("nerf_generate_rays_from_pixels", P(pixel_indices, torch.int64), int(first_pixel), n, cam, P(ray_o),
                   P(ray_d), st)
        return ray_o, ray_d, n

    @torch.no_grad()
    def render_rays(self, ray_o, ray_d, near: float, far: float, uniforms=None, want_depth: bool = False):
        """Coarse + fine forward for given rays.  uniforms = (u_c, u0, u1, u2) or None.  `want_depth` adds the fine pass's
        per-ray depth and opacity to the result."""
        n = ray_o.shape[0]
        with torch.cuda.device(self.device):
            uc = self._uniforms(n, None if uniforms is None else uniforms[:1], fine=False)
            co = self._pass(0, "c", ray_o, ray_d, n, float(near), float(far), None, uc, train=False)
            uf = self._uniforms(n, None if uniforms is None else uniforms[1:], fine=True)
            # the fine pass perturbs ITS copy of the coarse weights in place (utils.py:31); the returned coarse
            # weights stay as rendered
            w_pdf = self._get("w_pdf", (n, self.sc))
            w_pdf.copy_(co["w"])
            fi = self._pass(1, "f", ray_o, ray_d, n, float(near), float(far), w_pdf, uf, train=False, want_depth=want_depth)
        self.last = {"coarse": co, "fine": fi}
        out = {"rgb_coarse": co["rgb"], "weights_coarse": co["w"], "rgb_fine": fi["rgb"], "weights_fine": fi["w"],
               "t_fine": fi["t"]}
        if want_depth:
            out["depth_fine"], out["opacity_fine"] = fi["depth"], fi["opacity"]
        return out

    @torch.no_grad()
    def render_frame(self, camera: PerspectiveCamera, project_to_ndc: bool = False, first_pixel: int = 0,
                     count: Optional[int] = None, max_rays: int = 1 << 20, uniforms=None) -> torch.Tensor:
        """runners/render.py:58-107: all (or a contiguous range of) pixels, coarse then fine, clamped to [0,1].
        Returns (count, 3); reshape to (H, W, 3) for a full frame.  Rays are processed in chunks of at most `max_rays`
        (the reference's num_ray_batch, volume_renderer.py:229-254; ~7 KB of workspace per ray, so the default keeps a
        chunk near 7 GB) -- an 800x800 or 1008x756 frame is one chunk.  `uniforms` = (u_c, u0, u1, u2) over all `count`
        rays replays given draws (tests); otherwise each chunk draws its own in the reference's order."""
        total = camera.img_height * camera.img_width
        count = total - first_pixel if count is None else count
        max_rays = max(1, int(max_rays))
        if count <= max_rays:
            with torch.cuda.device(self.device):
                ray_o, ray_d, n = self.rays_from_pixels(camera, project_to_ndc, None, first_pixel, count)
                out = self.render_rays(ray_o, ray_d, camera.t_near, camera.t_far, uniforms)
            return out["rgb_fine"].clamp(0.0, 1.0)
        img = torch.empty((count, 3), device=self.device, dtype=torch.float32)
        for lo in range(0, count, max_rays):
            k = min(max_rays, count - lo)
            with torch.cuda.device(self.device):
                ray_o, ray_d, n = self.rays_from_pixels(camera, project_to_ndc, None, first_pixel + lo, k)
                u = None if uniforms is None else tuple(x[lo:lo + k] for x in uniforms)
                out = self.render_rays(ray_o, ray_d, camera.t_near, camera.t_far, u)
            torch.clamp(out["rgb_fine"], 0.0, 1.0, out=img[lo:lo + k])
        return img

    def train_rays(self, ray_o, ray_d, near: float, far: float, target: torch.Tensor, uniforms=None,
                   loss_out: Optional[torch.Tensor] = None):
        """Forward + backward of one training iteration (train.py:172-215) on given rays.  Gradients of both
        networks are OVERWRITTEN in the flat gradient buffer; returns the device tensor [coarse_loss, fine_loss].

        bf16 mode keeps the launch count down (the iteration is ~1 % small kernels): ONE prologue launch packs both
        networks' bf16 images and zeroes the gradients and the loss accumulators, the four uniform tensors come from one
        draw, and the MSE head rides on the compositing backward."""
        P, st = _lib.ptr, _lib.stream()
        flat = self.enable_flat_params()
        n = ray_o.shape[0]
        near, far = float(near), float(far)
        bf16 = self.precision == "bf16"
        with torch.cuda.device(self.device), torch.no_grad():
            losses = loss_out if loss_out is not None else self._get("losses", (2,))
            if bf16:
                c, f = self.nets
                for net in self.nets:
                    if net._packed is None or net._packed.device != self.device:
                        net._packed = torch.empty((self.lib.nerf_mlp_bf16_packed_bytes(),), device=self.device, dtype=torch.uint8)
                self._call("nerf_train_prologue", _lib.pointer_array([p.detach() for p in c.ordered_parameters()]),
                           P(c._packed, torch.uint8), _lib.pointer_array([p.detach() for p in f.ordered_parameters()]),
                           P(f._packed, torch.uint8), P(flat.grad), flat.grad.numel(), P(losses), 2, st)
                for net in self.nets:
                    net._packed_version = None  # packed outside NeRF.packed_weights' version tracking
                self._prologue_done = True
            else:
                losses.zero_()
            try:
                if uniforms is None:
                    # one draw for the four tensors of the reference's order (u_c | u0 | u1 | u2), each contiguous
                    sc, sf = self.sc, self.sf
                    u_all = torch.rand((n * (2 * sc + 2 * sf),), device=self.device)
                    o1, o2, o3 = n * sc, 2 * n * sc, 2 * n * sc + n * sf
                    uc = [u_all[:o1].view(n, sc)]
                    uf = [u_all[o1:o2].view(n, sc), u_all[o2:o3].view(n, sf), u_all[o3:].view(n, sf)]
                else:
                    uc = self._uniforms(n, uniforms[:1], fine=False)
                    uf = self._uniforms(n, uniforms[1:], fine=True)
                co = self._pass(0, "c", ray_o, ray_d, n, near, far, None, uc, train=True)
                # the fine pass perturbs the coarse weights in place (utils.py:31); nothing reads them afterwards (the
                # compositing backward recomputes the weights), so no private copy is made here
                fi = self._pass(1, "f", ray_o, ray_d, n, near, far, co["w"], uf, train=True)
                # autograd order: the fine pass is differentiated first (train.py:190-215)
                self._backward(1, fi, n, None, flat.grads_of(1), target=target, loss=losses[1:2])
                self._backward(0, co, n, None, flat.grads_of(0), target=target, loss=losses[0:1])
            finally:
                self._prologue_done = False
        self.last = {"coarse": co, "fine": fi}
        return losses

    def train_pixels(self, camera: PerspectiveCamera, pixel_indices: torch.Tensor, target: torch.Tensor,
                     project_to_ndc: bool = False, uniforms=None, loss_out=None):
        with torch.cuda.device(self.device):
            ray_o, ray_d, n = self.rays_from_pixels(camera, project_to_ndc, pixel_indices)
        return self.train_rays(ray_o, ray_d, camera.t_near, camera.t_far, target, uniforms, loss_out)

    # ------------------------------------------------------------------------------------------ captured iteration
    def _train_static(self, sg: "_StepGraph"):
        """The iteration on the graph's static inputs (pixel ids, targets and the camera all live in device memory)."""
        P, st = _lib.ptr, _lib.stream()
        ray_o = self._get("ray_o", (sg.n, 3))
        ray_d = self._get("ray_d", (sg.n, 3))
        self._call("nerf_generate_rays_from_pixels_devcam", P(sg.pix, torch.int64), 0, sg.n, P(sg.cam, torch.uint8), P(ray_o),
                   P(ray_d), st)
        self.train_rays(ray_o, ray_d, sg.near, sg.far, sg.tgt, None, loss_out=sg.losses)

    def train_pixels_graph(self, camera: PerspectiveCamera, pixel_indices: torch.Tensor, target: torch.Tensor,
                           project_to_ndc: bool = False) -> torch.Tensor:
        """`train_pixels` replayed from a CUDA graph: per iteration the host enqueues two input copies (from pinned host
        or device memory), one camera upload and ONE graph launch instead of ~45 kernel launches and a dozen torch ops.

        The reference builds a new camera and a new pixel batch every iteration (train.py:136-171), so everything that
        changes per iteration is read from device memory by the captured kernels: pixel ids and targets from static
        buffers, the camera from a device struct (nerf_upload_camera), the uniform draws from torch's graph-safe Philox
        state.  Scene bounds, ray count and NDC flag are launch constants: one graph per (n, near, far, ndc).
        The first call for a key runs the iteration eagerly (that IS this iteration's result; it also sizes every
        workspace) and then captures it; later calls replay.  Returns the device tensor [coarse_loss, fine_loss]."""
        n = int(pixel_indices.shape[0])
        key = (n, float(camera.t_near), float(camera.t_far), bool(project_to_ndc))
        sg = self._graphs.get(key)
        fresh = sg is None
        with torch.cuda.device(self.device):
            if fresh:
                sg = _StepGraph(n, key[1], key[2], self.device)
                self._graphs[key] = sg
            sg.pix.copy_(pixel_indices, non_blocking=True)
            sg.tgt.copy_(target, non_blocking=True)
            self._call("nerf_upload_camera", _lib.ptr(sg.cam, torch.uint8), camera.pack(project_to_ndc), _lib.stream())
            if fresh:
                self._train_static(sg)
                torch.cuda.synchronize(self.device)
                l0 = self.launches
                graph = torch.cuda.CUDAGraph()
                try:
                    with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                        self._train_static(sg)
                    sg.graph = graph
                    sg.launches = self.launches - l0
                    sg.keepalive = list(self._buf.values())
                except RuntimeError as err:  # capture refused (e.g. another library touched the stream): stay eager
                    import warnings

                    warnings.warn(f"train_pixels_graph: CUDA graph capture failed, running eagerly ({err})")
                    sg.graph = None
                self.launches = l0
            elif sg.graph is not None:
                sg.graph.replay()
                self.launches += sg.launches
            else:
                self._train_static(sg)
        return sg.losses
