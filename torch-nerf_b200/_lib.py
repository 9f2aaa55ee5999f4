"""ctypes binding of libnerf_b200.so (include/nerf_b200.h).

There is NO fallback: if the shared library is missing the import of any compute entry point raises, and
every wrapper refuses tensors that are not CUDA tensors.  torch is used here only for device memory and the
current stream.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
# NERF_B200_LIB_SUFFIX selects an experimental build made by build.py with the same variable (kernel variants under test)
_SUFFIX = os.environ.get("NERF_B200_LIB_SUFFIX", "")
LIB_PATH = os.path.join(_PKG, "lib", f"libnerf_b200{_SUFFIX}.so")
SELFTEST_LIB_PATH = os.path.join(_PKG, "lib", f"libnerf_b200{_SUFFIX}_selftest.so")

NERF_OK, NERF_ERR_ARG, NERF_ERR_CUDA = 0, 1, 2
NUM_PARAM_TENSORS = 22


class CameraStruct(ctypes.Structure):
    """nerf_camera_t"""

    _fields_ = [
        ("fx", c_float), ("fy", c_float), ("cx", c_float), ("cy", c_float),
        ("rot", c_float * 9), ("trans", c_float * 3),
        ("ndc_sx", c_float), ("ndc_sy", c_float), ("ndc_two_near", c_float),
        ("img_w", c_int32), ("img_h", c_int32), ("project_to_ndc", c_int32),
    ]


class MlpDims(ctypes.Structure):
    """nerf_mlp_dims_t"""

    _fields_ = [("pos_dim", c_int32), ("view_dim", c_int32), ("feat_dim", c_int32)]


_P = c_void_p
_PROTOTYPES = {
    # name: (restype, argtypes)
    "nerf_version": (c_int, []),
    "nerf_last_error": (c_char_p, []),
    "nerf_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "nerf_generate_rays": (c_int, [_P, c_int64, POINTER(CameraStruct), _P, _P, _P]),
    "nerf_generate_rays_from_pixels": (c_int, [_P, c_int64, c_int64, POINTER(CameraStruct), _P, _P, _P]),
    "nerf_map_rays_to_ndc": (c_int, [_P, _P, c_int64, c_double, c_double, c_int, c_int, _P, _P, _P]),
    "nerf_upload_camera": (c_int, [_P, POINTER(CameraStruct), _P]),
    "nerf_generate_rays_from_pixels_devcam": (c_int, [_P, c_int64, c_int64, _P, _P, _P, _P]),
    "nerf_make_bins": (c_int, [c_double, c_double, c_int, POINTER(c_float), POINTER(c_float)]),
    "nerf_sample_coarse": (c_int, [_P, _P, c_int64, c_int, c_double, c_double, _P, _P, _P, _P, _P, _P]),
    "nerf_sample_pdf": (c_int, [c_double, c_double, _P, _P, _P, c_int64, c_int, c_int, _P, _P, _P]),
    "nerf_sample_fine": (c_int, [_P, _P, c_int64, c_int, c_int, c_double, c_double, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "nerf_posenc": (c_int, [_P, c_int64, c_int, c_int, c_int, _P, c_int64, _P]),
    "nerf_composite_fwd": (c_int, [_P, _P, _P, _P, c_int64, c_int, _P, _P, _P, _P, _P]),
    "nerf_composite_bwd": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int, _P, _P, _P]),
    "nerf_composite_bwd_mse": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int, _P, _P, _P, _P]),
    "nerf_train_prologue": (c_int, [POINTER(_P), _P, POINTER(_P), _P, _P, c_int64, _P, c_int64, _P]),
    "nerf_mse_loss": (c_int, [_P, _P, c_int64, _P, _P, _P]),
    "nerf_adam_step": (c_int, [_P, _P, _P, _P, c_int64, c_double, c_double, c_double, c_double, c_int64, c_double, _P]),
    "nerf_dp_exchange_adam": (c_int, [POINTER(_P), POINTER(_P), c_int, c_int, _P, _P, _P, c_int64, c_double, c_double, c_double,
                                      c_double, c_int64, c_double, ctypes.c_uint32, _P, _P]),
    "nerf_mlp_f32_cache_floats": (c_size_t, [POINTER(MlpDims), c_int64]),
    "nerf_mlp_f32_bwd_scratch_floats": (c_size_t, [POINTER(MlpDims), c_int64]),
    "nerf_mlp_f32_forward": (c_int, [POINTER(MlpDims), POINTER(_P), _P, _P, c_int64, _P, _P, _P, _P]),
    "nerf_mlp_f32_backward": (c_int, [POINTER(MlpDims), POINTER(_P), _P, _P, c_int64, _P, _P, POINTER(_P), _P, _P]),
    "nerf_mlp_bf16_packed_bytes": (c_size_t, []),
    "nerf_mlp_bf16_pack": (c_int, [POINTER(_P), _P, _P]),
    "nerf_mlp_bf16_cache_bytes": (c_size_t, [c_int64]),
    "nerf_mlp_bf16_forward": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int64, _P, _P, _P, _P]),
    "nerf_mlp_bf16_bwd_scratch_bytes": (c_size_t, [c_int64]),
    "nerf_mlp_bf16_backward": (c_int, [_P, _P, _P, c_int64, _P, _P, POINTER(_P), _P, _P]),
    "nerf_mlp_bf16_backward_part": (c_int, [_P, _P, _P, c_int64, _P, _P, POINTER(_P), _P, c_int, c_int64, c_int64, c_int, _P]),
}
# include/nerf_b200_debug.h: timeline hooks of the main library ...
_DEBUG_PROTOTYPES = {
    "nerf_debug_set_profile_buffer": (c_int, [_P, c_int]),
    "nerf_debug_set_wgrad_profile": (c_int, [_P]),
    "nerf_debug_set_wgrad_wrap": (c_int, [c_int64, c_int64]),
    "nerf_debug_set_wgrad_costs": (c_int, [POINTER(c_int), c_int]),
}
# ... and the building-block self tests / micro-benchmarks of libnerf_b200_selftest.so (tests/ and tools/ only)
_SELFTEST_PROTOTYPES = {
    "nerf_selftest_umma": (c_int, [_P, _P, _P, c_int, c_int, c_int, _P]),
    "nerf_selftest_mma_rate": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "nerf_selftest_umma2": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "nerf_selftest_write_bw": (c_int, [_P, c_size_t, c_int, c_int, _P]),
}

EXPORTED_SYMBOLS = tuple(_PROTOTYPES)
DEBUG_SYMBOLS = tuple(_DEBUG_PROTOTYPES)
SELFTEST_SYMBOLS = tuple(_SELFTEST_PROTOTYPES)

_lib = None
_selftest_lib = None


def load() -> ctypes.CDLL:
    """Loads the library once; raises (never falls back) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python torch-nerf_b200/build.py` "
                "(there is no CPU or PyTorch fallback for the B200 path)"
            )
        lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        for name, (res, args) in {**_PROTOTYPES, **_DEBUG_PROTOTYPES}.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def load_selftest() -> ctypes.CDLL:
    """libnerf_b200_selftest.so: the tcgen05 building-block checks and micro-benchmarks (tests/ and tools/ only)."""
    global _selftest_lib
    if _selftest_lib is None:
        load()
        if not os.path.exists(SELFTEST_LIB_PATH):
            raise RuntimeError(f"{SELFTEST_LIB_PATH} is missing: build it with `python torch-nerf_b200/build.py`")
        lib = ctypes.CDLL(SELFTEST_LIB_PATH)
        for name, (res, args) in _SELFTEST_PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _selftest_lib = lib
    return _selftest_lib


def last_error() -> str:
    msg = load().nerf_last_error()
    return msg.decode() if msg else ""


def check(rc: int, what: str) -> None:
    if rc == NERF_OK:
        return
    msg = f"{what}: {last_error()}"
    if rc == NERF_ERR_ARG:
        raise ValueError(msg)
    raise RuntimeError(msg)


def ptr(t: torch.Tensor | None, dtype=torch.float32):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("libnerf_b200 operates on CUDA tensors only (no CPU fallback)")
    if t.dtype != dtype:
        raise ValueError(f"expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError("expected a contiguous tensor")
    return c_void_p(t.data_ptr())


def stream() -> c_void_p:
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def pointer_array(tensors) -> ctypes.Array:
    arr = (_P * len(tensors))()
    for i, t in enumerate(tensors):
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError("parameter tensors must be contiguous float32 CUDA tensors")
        arr[i] = t.data_ptr()
    return arr
