"""
CPU ORACLE (bf16 emulation) -- TEST INFRASTRUCTURE ONLY.

The fp32 oracle (nerf_oracle.py) restates the reference; this module restates the SAME network
(network/nerf.py:65-121) with the rounding points of the tensor-core kernels made explicit, so that a kernel
bug can be told apart from bf16 quantisation:
  * every tensor-core operand (layer inputs, weights, activation gradients) is rounded to bfloat16
    (round-to-nearest-even), products are exact and accumulation is float32/float64;
  * biases, the density head (fc_8 row 0), fc_out and the sigmoid stay float32, computed from the UNROUNDED
    float32 layer outputs, exactly like the kernels' epilogues;
  * ReLU masks are taken from the float32 layer outputs.
Parity status: derived from the pinned fp32 oracle (same formulas); agreement with it is bounded by bf16
resolution and checked in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import numpy as np

from . import nerf_oracle as orc

F32 = np.float32


def bf16_round(x: np.ndarray) -> np.ndarray:
    """float32 -> bfloat16 (round to nearest even) -> float32."""
    x = np.ascontiguousarray(x, dtype=F32)
    u = x.view(np.uint32).astype(np.uint64)
    rounded = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return rounded.astype(np.uint32).view(F32).reshape(x.shape)


def _mm(a, b):
    return (a.astype(np.float64) @ b.astype(np.float64)).astype(F32)


def forward(params: dict, pts: np.ndarray, dirs: np.ndarray):
    """Returns sigma (M,), rgb (M,3) and the cache the backward needs."""
    W = {k: v for k, v in params.items()}
    Wb = {k: bf16_round(v) for k, v in params.items() if k.endswith("weight")}
    relu = lambda v: np.maximum(v, F32(0))
    pe_b = bf16_round(orc.positional_encode(pts, 10))
    de_b = bf16_round(orc.positional_encode(dirs, 4))
    c = {"pe_b": pe_b, "de_b": de_b}
    x_b = pe_b
    names = ["fc_in", "fc_1", "fc_2", "fc_3", "fc_4", "fc_5", "fc_6", "fc_7"]
    for l, name in enumerate(names):
        if l == 5:
            x_b = np.concatenate([pe_b, x_b], axis=-1)
        h = relu(_mm(x_b, Wb[f"{name}.weight"].T) + W[f"{name}.bias"])
        c[f"h{l}"] = h
        c[f"h{l}_b"] = bf16_round(h)
        x_b = c[f"h{l}_b"]
    h7 = c["h7"]
    sigma_pre = (h7.astype(np.float64) @ W["fc_8.weight"][0].astype(np.float64)).astype(F32) + W["fc_8.bias"][0]
    feat = _mm(c["h7_b"], Wb["fc_8.weight"][1:].T) + W["fc_8.bias"][1:]
    c["feat_b"] = bf16_round(feat)
    x9_b = np.concatenate([c["feat_b"], de_b], axis=-1)
    h9 = relu(_mm(x9_b, Wb["fc_9.weight"].T) + W["fc_9.bias"])
    c["h9"], c["h9_b"] = h9, bf16_round(h9)
    z = _mm(h9, W["fc_out.weight"].T) + W["fc_out.bias"]
    rgb = (F32(1) / (F32(1) + np.exp(-z))).astype(F32)
    c["rgb"], c["sigma_pre"] = rgb, sigma_pre
    return relu(sigma_pre), rgb, c


def backward(params: dict, c: dict, g_sigma: np.ndarray, g_rgb: np.ndarray) -> dict:
    W = params
    Wb = {k: bf16_round(v) for k, v in params.items() if k.endswith("weight")}
    g = {}
    rgb = c["rgb"]
    gz = (g_rgb * rgb * (F32(1) - rgb)).astype(F32)
    gsp = (g_sigma * (c["sigma_pre"] > 0)).astype(F32)
    g["fc_out.weight"] = _mm(gz.T, c["h9_b"])
    g["fc_out.bias"] = gz.sum(0).astype(F32)
    G9_b = bf16_round(_mm(gz, W["fc_out.weight"]) * (c["h9"] > 0))
    x9_b = np.concatenate([c["feat_b"], c["de_b"]], axis=-1)
    g["fc_9.weight"] = _mm(G9_b.T, x9_b)
    g["fc_9.bias"] = G9_b.sum(0).astype(F32)
    G8f_b = bf16_round(_mm(G9_b, Wb["fc_9.weight"][:, :256]))
    g["fc_8.weight"] = np.concatenate([_mm(gsp[None, :], c["h7_b"]), _mm(G8f_b.T, c["h7_b"])], axis=0)
    g["fc_8.bias"] = np.concatenate([[gsp.sum()], G8f_b.sum(0)]).astype(F32)
    d = _mm(G8f_b, Wb["fc_8.weight"][1:]) + gsp[:, None] * W["fc_8.weight"][0][None, :]
    G_b = bf16_round(d * (c["h7"] > 0))  # G7
    for l in (7, 6):
        name = f"fc_{l}"
        g[f"{name}.weight"] = _mm(G_b.T, c[f"h{l - 1}_b"])
        g[f"{name}.bias"] = G_b.sum(0).astype(F32)
        G_b = bf16_round(_mm(G_b, Wb[f"{name}.weight"]) * (c[f"h{l - 1}"] > 0))
    # G_b = G5
    x5_b = np.concatenate([c["pe_b"], c["h4_b"]], axis=-1)
    g["fc_5.weight"] = _mm(G_b.T, x5_b)
    g["fc_5.bias"] = G_b.sum(0).astype(F32)
    G_b = bf16_round(_mm(G_b, Wb["fc_5.weight"][:, 63:]) * (c["h4"] > 0))  # G4
    for l in (4, 3, 2, 1):
        name = f"fc_{l}"
        g[f"{name}.weight"] = _mm(G_b.T, c[f"h{l - 1}_b"])
        g[f"{name}.bias"] = G_b.sum(0).astype(F32)
        G_b = bf16_round(_mm(G_b, Wb[f"{name}.weight"]) * (c[f"h{l - 1}"] > 0))
    g["fc_in.weight"] = _mm(G_b.T, c["pe_b"])
    g["fc_in.bias"] = G_b.sum(0).astype(F32)
    return g
