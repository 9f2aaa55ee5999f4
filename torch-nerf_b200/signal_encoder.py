"""Positional encoder backed by `nerf_posenc`.  Mirror of `SignalEncoderBase` / `PositionalEncoder`
(reference src/signal_encoder/signal_encoder_base.py:8-28, positional_encoder.py:12-114)."""
from __future__ import annotations

import torch

from . import _lib


class SignalEncoderBase:
    def __init__(self):
        pass

    def encode(self, in_signal: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError()


class PositionalEncoder(SignalEncoderBase):
    def __init__(self, in_dim: int, embed_level: int, include_input: bool):
        super().__init__()
        self._embed_level = embed_level
        self._include_input = include_input
        self._in_dim = in_dim
        self._out_dim = 2 * self._embed_level * self._in_dim + (self._in_dim if include_input else 0)

    def encode(self, in_signal: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        """(N, C) -> (N, out_dim): [x | sin(2^0 x) | cos(2^0 x) | ... ] (positional_encoder.py:83-104)."""
        lib = _lib.load()
        if not in_signal.is_cuda:
            raise RuntimeError("torch_nerf_b200 runs on CUDA tensors only (no CPU fallback)")
        x = in_signal.detach().to(torch.float32).contiguous()
        m = x.shape[0]
        if out is None:
            out = torch.empty((m, self._out_dim), device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(
                lib.nerf_posenc(_lib.ptr(x), m, self._in_dim, self._embed_level, 1 if self._include_input else 0,
                                _lib.ptr(out), out.stride(0), _lib.stream()),
                "nerf_posenc",
            )
        return out

    @property
    def in_dim(self) -> int:
        return self._in_dim

    @property
    def out_dim(self) -> int:
        return self._out_dim

    @property
    def embed_level(self) -> int:
        return self._embed_level

    @property
    def include_input(self) -> bool:
        return self._include_input
