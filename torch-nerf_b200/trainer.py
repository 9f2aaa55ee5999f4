"""Callers of the hot path: the reference's training iteration and its full-frame
render / validation step, restated over the fused engine.

    train_one_epoch   ref runners/train.py:105-218   (centre-crop warm-up, coarse + fine MSE, Adam, ExponentialLR)
    optimizer setup   ref runners/runner_utils.py:663-717
    render_image      ref runners/runner_utils.py:834-918, runners/train.py:296-341 (clamp, (H*W,3) -> (3,H,W), PNG)

Differences that are deliberate: both networks live in one flat buffer stepped by ONE Adam launch (optim.FlatAdam); the loss is
accumulated on the device and read back once per epoch (the reference calls .item() three times per iteration); with
several ranks every rank trains its shard of the pixel batch and the flat gradient is all-reduced once per step.
By default (`use_graph=True`, bf16) the forward + backward of an iteration is replayed from a CUDA graph
(engine.train_pixels_graph); the all-reduce and the one-launch Adam stay outside it.  `use_graph=False` (and fp32
validation mode) enqueues every kernel of the iteration instead.  With several ranks the constructor and `load_ckpt`
broadcast rank 0's parameters and optimizer state, so replicas cannot drift apart through different seeds."""
from __future__ import annotations

from typing import Dict, Iterable, Tuple

import torch

from .cameras import PerspectiveCamera
from .checkpoint import load_ckpt, save_ckpt
from .network import NeRF
from .parallel import allreduce_mean_, shard_range


def center_crop_pixel_indices(img_height: int, img_width: int) -> torch.Tensor:
    """train.py:151-167: flat ids (row * W + col) of the central block the first 10 epochs sample from.
    rows [ci - ci//2, ci + ci//2), cols [cj - cj//2, cj + cj//2) with ci = (H-1)//2, cj = (W-1)//2, row-major."""
    ci, cj = (img_height - 1) // 2, (img_width - 1) // 2
    rows = torch.arange(ci - ci // 2, ci + ci // 2)
    cols = torch.arange(cj - cj // 2, cj + cj // 2)
    return (rows[:, None] * img_width + cols[None, :]).reshape(-1)


def exp_lr_gamma(init_lr: float, end_lr: float, num_iter: int) -> float:
    """runner_utils.py:704-708"""
    return pow(end_lr / init_lr, 1 / num_iter)


def psnr(pred: torch.Tensor, target: torch.Tensor, data_range: float = 1.0) -> torch.Tensor:
    """10 log10(range^2 / MSE), on the device (stands in for torchmetrics' PeakSignalNoiseRatio, train.py:349)."""
    mse = torch.mean((pred.to(torch.float32) - target.to(torch.float32)) ** 2)
    return 10.0 * torch.log10(torch.as_tensor(data_range ** 2, device=mse.device) / mse)


def save_png(img_chw: torch.Tensor, path: str) -> None:
    """torchvision.utils.save_image for one image (runner_utils.py:913-917): x*255 + 0.5, clamp, uint8, PNG."""
    from PIL import Image

    arr = img_chw.detach().to(torch.float32).mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to("cpu", torch.uint8).numpy()
    Image.fromarray(arr).save(path)


class Trainer:
    """The reference's training loop state: two networks, Adam(lr=init_lr, eps) and a per-iteration ExponentialLR."""

    def __init__(self, coarse: NeRF, fine: NeRF, num_samples_coarse: int = 64, num_samples_fine: int = 128,
                 num_pixels: int = 4096, t_near: float = 2.0, t_far: float = 6.0, project_to_ndc: bool = False,
                 init_lr: float = 5e-4, end_lr: float = 5e-5, num_iter: int = 300000, eps: float = 1e-8,
                 precision: str = "bf16", rank: int = 0, world: int = 1, seed: int = 0, use_graph: bool = True):
        from .engine import HotPathEngine

        self.use_graph = bool(use_graph) and precision == "bf16"

        self.coarse, self.fine = coarse, fine
        self.engine = HotPathEngine(coarse, fine, num_samples_coarse, num_samples_fine, precision)
        # several ranks: the flat gradient buffer lives in symmetric memory and the exchange is fused with Adam
        # (parallel.PeerExchange); None (single rank, or no symmetric memory) -> NCCL all-reduce + one Adam launch
        from .parallel import make_exchange

        numel = sum(p.numel() for net in (coarse, fine) for p in net.parameters())
        self.exchange = make_exchange(numel, self.engine.device, int(world))
        self.flat = self.engine.enable_flat_params(self.exchange.grad if self.exchange is not None else None)
        self.num_pixels = int(num_pixels)
        self.t_near, self.t_far, self.project_to_ndc = float(t_near), float(t_far), bool(project_to_ndc)
        self.rank, self.world = int(rank), int(world)
        from .optim import FlatAdam

        self.optimizer = FlatAdam([self.flat.param], lr=init_lr, eps=eps)  # torch.optim.Adam's state layout, one launch
        self.optimizer.grad_scale = 1.0 / self.world
        self.optimizer.exchange = self.exchange
        self.scheduler = torch.optim.lr_scheduler.ExponentialLR(self.optimizer, exp_lr_gamma(init_lr, end_lr, num_iter))
        self.device = self.engine.device
        # pixel selection is host-side in the reference (np.random.choice / torch.randperm on CPU): one generator,
        # identical on every rank, so the ranks agree on the global batch they shard
        self._gen = torch.Generator().manual_seed(seed)
        self._losses = torch.zeros(2, device=self.device)
        self.sync_replicas()

    def sync_replicas(self) -> None:
        """Data-parallel replicas must hold identical parameters and Adam state: rank 0's are broadcast (a caller that
        seeds the ranks differently, or ranks that read different checkpoint files, would otherwise train different
        weights on averaged gradients without any error)."""
        if self.world > 1:
            from .parallel import broadcast_replica_state

            broadcast_replica_state(self.flat.param, self.optimizer, self.scheduler)

    # ------------------------------------------------------------------------------------------ training
    def select_pixels(self, img_height: int, img_width: int, epoch: int) -> torch.Tensor:
        """Global pixel batch of one iteration: centre block while epoch < 10 (train.py:148-167), otherwise a
        no-replacement draw over the frame (volume_renderer.py:139-146)."""
        n = self.num_pixels * self.world
        if epoch < 10:
            cand = center_crop_pixel_indices(img_height, img_width)
            return cand[torch.randperm(len(cand), generator=self._gen)[:n]]
        return torch.randperm(img_height * img_width, generator=self._gen)[:n]

    def train_iteration(self, pixel_gt: torch.Tensor, camera: PerspectiveCamera, epoch: int) -> torch.Tensor:
        """One pass of the loop body of train.py:120-213.  pixel_gt: (H*W, 3) ground-truth colours (host or device).
        Returns the device tensor [coarse_loss, fine_loss] of this rank's shard (no synchronisation)."""
        pix = self.select_pixels(camera.img_height, camera.img_width, epoch)
        lo, hi = shard_range(len(pix), self.rank, self.world)
        pix = pix[lo:hi]
        tgt = pixel_gt.reshape(-1, 3)[pix.to(pixel_gt.device)].to(torch.float32).contiguous()
        if self.use_graph:
            # the whole forward + backward replayed from one CUDA graph; pixel ids and colours are
            # copied straight into the graph's static buffers
            losses = self.engine.train_pixels_graph(camera, pix, tgt, self.project_to_ndc)
        else:
            losses = self.engine.train_pixels(camera, pix.to(self.device, non_blocking=True),
                                              tgt.to(self.device, non_blocking=True), self.project_to_ndc,
                                              loss_out=self._losses)
        if self.exchange is None:
            allreduce_mean_(self.flat.grad, self.world, scale=False)  # the 1/world rides on the Adam kernel
        self.optimizer.step()  # with a PeerExchange: gradient sum over the ranks + Adam in one launch
        self.scheduler.step()
        return losses

    def train_one_epoch(self, batches: Iterable[Tuple[torch.Tensor, torch.Tensor]], intrinsic: Dict[str, float],
                        epoch: int) -> Dict[str, float]:
        """train.py:105-218.  `batches` yields (pixel_gt (H, W, 3) or (H*W, 3), extrinsic (4,4) / (3,4)) like the
        reference's DataLoader with batch size 1; `intrinsic` = {"f_x", "f_y", "img_width", "img_height"}.
        Returns {"coarse_loss", "fine_loss", "loss"} averaged over the batches (one device->host read)."""
        total = torch.zeros(2, device=self.device)
        count = 0
        for pixel_gt, extrinsic in batches:
            camera = PerspectiveCamera(intrinsic, torch.as_tensor(extrinsic).squeeze(), self.t_near, self.t_far)
            total += self.train_iteration(torch.as_tensor(pixel_gt).squeeze().reshape(-1, 3), camera, epoch)
            count += 1
        c, f = (total / max(count, 1)).tolist()
        return {"coarse_loss": c, "fine_loss": f, "loss": c + f}

    # ------------------------------------------------------------------------------------------ rendering
    @torch.no_grad()
    def render_image(self, camera: PerspectiveCamera) -> torch.Tensor:
        """Full frame through both networks, clamped to [0, 1], as (3, H, W) (train.py:330-341).  With several ranks
        every rank renders its contiguous pixel range and the (H*W, 3) image is all-gathered."""
        h, w = camera.img_height, camera.img_width
        lo, hi = shard_range(h * w, self.rank, self.world)
        part = self.engine.render_frame(camera, self.project_to_ndc, lo, hi - lo)
        if self.world > 1:
            import torch.distributed as dist

            sizes = [shard_range(h * w, r, self.world) for r in range(self.world)]
            parts = [torch.empty((e - s, 3), device=self.device) for s, e in sizes]
            dist.all_gather(parts, part.contiguous())
            part = torch.cat(parts, 0)
        return part.reshape(h, w, 3).permute(2, 0, 1)

    @torch.no_grad()
    def validate(self, pixel_gt: torch.Tensor, camera: PerspectiveCamera) -> Tuple[torch.Tensor, torch.Tensor]:
        """train.py:296-352 for one view: returns (image (3,H,W), PSNR) with both images clamped to [0,1]."""
        img = self.render_image(camera)
        gt = torch.as_tensor(pixel_gt).reshape(camera.img_height, camera.img_width, 3).permute(2, 0, 1).to(self.device).clamp(0.0, 1.0)
        return img, psnr(img, gt)

    # ------------------------------------------------------------------------------------------ checkpoints
    def save_ckpt(self, ckpt_dir, epoch: int) -> str:
        return save_ckpt(ckpt_dir, epoch, self.coarse, self.fine, self.optimizer, self.scheduler, self.flat)

    def load_ckpt(self, ckpt_dir) -> int:
        epoch = load_ckpt(ckpt_dir, self.coarse, self.fine, self.optimizer, self.scheduler, self.flat)
        self.sync_replicas()
        return epoch
