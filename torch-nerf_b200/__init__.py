"""torch_nerf_b200: B200-native (sm_100a) implementation of torch-NeRF's per-ray rendering hot path.

The classes below keep the reference's call signatures and forward to the C-ABI library
libnerf_b200.so (include/nerf_b200.h).  There is no CPU fallback."""
from . import _lib
from .cameras import PerspectiveCamera
from .integrators import IntegratorBase, QuadratureIntegrator
from .network import NeRF
from .ray_samplers import RayBundle, RaySamplerBase, StratifiedSampler, make_bins, sample_pdf
from .scene import PrimitiveBase, PrimitiveCube
from .signal_encoder import PositionalEncoder, SignalEncoderBase
from .volume_renderer import VolumeRenderer
from . import checkpoint, datasets
from .datasets import BlenderDataset, LLFFDataset
from .optim import FlatAdam
from .trainer import Trainer, center_crop_pixel_indices, exp_lr_gamma

__all__ = [
    "PerspectiveCamera", "IntegratorBase", "QuadratureIntegrator", "NeRF", "RayBundle", "RaySamplerBase",
    "StratifiedSampler", "make_bins", "sample_pdf", "PrimitiveBase", "PrimitiveCube", "PositionalEncoder",
    "SignalEncoderBase", "VolumeRenderer", "Trainer", "FlatAdam", "checkpoint", "datasets",
    "BlenderDataset", "LLFFDataset", "center_crop_pixel_indices", "exp_lr_gamma",
]
