"""Scene primitives.  Mirror of `PrimitiveBase` / `PrimitiveCube` (reference src/scene/primitives/primitive_base.py:12-74,
cube.py:13-81): flatten (N,S,3), encode, query the network, reshape back.  When the network runs in bf16 mode and
the encoders are the stock (3,10,True)/(3,4,True) pair, encoding + MLP run as ONE tensor-core kernel."""
from __future__ import annotations

import warnings
from typing import Dict, Optional, Tuple

import torch

from .network import NeRF
from .signal_encoder import PositionalEncoder, SignalEncoderBase


class PrimitiveBase:
    def __init__(self, encoders: Optional[Dict[str, SignalEncoderBase]] = None):
        if encoders is not None:
            if not isinstance(encoders, dict):
                raise ValueError(f"Expected a parameter of type Dict. Got {type(encoders)}")
            if "coord_enc" not in encoders.keys():
                warnings.warn(f"Missing an encoder type 'coord_enc'. Got {encoders.keys()}.")
            if "dir_enc" not in encoders.keys():
                warnings.warn(f"Missing an encoder type 'dir_enc'. Got {encoders.keys()}.")
        self._encoders = encoders

    def query_points(self, pos: torch.Tensor, view_dir: torch.Tensor) -> Tuple[int, int]:
        if pos.shape != view_dir.shape:
            raise ValueError(f"Expected tensors of same shape. Got {pos.shape} and {view_dir.shape}, respectively.")
        num_ray, num_sample, _ = pos.shape
        return num_ray, num_sample

    @property
    def encoders(self) -> Optional[Dict[str, SignalEncoderBase]]:
        return self._encoders

    @encoders.setter
    def encoders(self, new_encoders) -> None:
        if not isinstance(new_encoders, dict):
            raise ValueError(f"Expected a parameter of type Dict. Got {type(new_encoders)}")
        if "coord_enc" not in new_encoders.keys():
            raise ValueError(f"Missing required encoder type 'coord_enc'. Got {new_encoders.keys()}.")
        if "dir_enc" not in new_encoders.keys():
            raise ValueError(f"Missing required encoder type 'dir_enc'. Got {new_encoders.keys()}.")
        self._encoders = new_encoders


def _is_stock_encoder(enc, level: int) -> bool:
    return isinstance(enc, PositionalEncoder) and enc.in_dim == 3 and enc.embed_level == level and enc.include_input


class PrimitiveCube(PrimitiveBase):
    def __init__(self, radiance_field: torch.nn.Module, encoders: Optional[Dict[str, SignalEncoderBase]] = None):
        super().__init__(encoders=encoders)
        if not isinstance(radiance_field, torch.nn.Module):
            raise ValueError(f"Expected a parameter of type torch.nn.Module. Got {type(radiance_field)}.")
        self._radiance_field = radiance_field

    def fused_bf16_available(self) -> bool:
        net, enc = self._radiance_field, self._encoders
        return (
            isinstance(net, NeRF) and net.precision == "bf16" and net.supports_bf16() and enc is not None
            and _is_stock_encoder(enc.get("coord_enc"), 10) and _is_stock_encoder(enc.get("dir_enc"), 4)
        )

    def query_points(self, pos: torch.Tensor, view_dir: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """pos, view_dir (N,S,3) -> sigma (N,S), radiance (N,S,3)   (cube.py:39-76)."""
        num_ray, num_sample = super().query_points(pos, view_dir)
        m = num_ray * num_sample
        if self.fused_bf16_available():
            sigma, radiance = self._radiance_field.query_raw(
                pos.reshape(m, 3).to(torch.float32).contiguous(), view_dir.reshape(m, 3).to(torch.float32).contiguous()
            )
            return sigma.reshape(num_ray, num_sample), radiance.reshape(num_ray, num_sample, -1)
        if self.encoders is not None:
            if "coord_enc" in self.encoders.keys():
                pos = self.encoders["coord_enc"].encode(pos.reshape(m, -1))
            if "dir_enc" in self.encoders.keys():
                view_dir = self.encoders["dir_enc"].encode(view_dir.reshape(m, -1))
        sigma, radiance = self._radiance_field(pos.reshape(m, -1), view_dir.reshape(m, -1))
        return sigma.reshape(num_ray, num_sample), radiance.reshape(num_ray, num_sample, -1)

    @property
    def radiance_field(self) -> torch.nn.Module:
        return self._radiance_field
