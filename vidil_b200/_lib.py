"""ctypes binding of the C ABI in include/vidil_b200.h.

The shared library is built in-tree by vidil_b200.build (nvcc, sm_100a).  There is no Python or CPU
fallback for any entry point: if the library is missing it is built, if it cannot be built or loaded
the import of the caller fails, and on a machine without a B200 every compute call raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int32, c_int64, c_size_t, c_void_p

from . import build as _build

ABI_VERSION = 2  # VIDIL_B200_ABI_VERSION of include/vidil_b200.h this binding was written against
DTYPE_BF16, DTYPE_FP16 = 0, 1
ACT_GELU_ERF, ACT_QUICK_GELU = 0, 1
EPI_STORE, EPI_GELU, EPI_QUICKGELU, EPI_RESID, EPI_PATCH, EPI_STORE_F32 = range(6)

DTYPES = {"bf16": DTYPE_BF16, "bfloat16": DTYPE_BF16, "fp16": DTYPE_FP16, "float16": DTYPE_FP16, "half": DTYPE_FP16}


class EncoderCfg(ctypes.Structure):
    """Mirror of `vidil_encoder_cfg` (include/vidil_b200.h)."""

    _fields_ = [
        ("img_size", c_int32),
        ("patch_size", c_int32),
        ("embed_dim", c_int32),
        ("depth", c_int32),
        ("num_heads", c_int32),
        ("mlp_dim", c_int32),
        ("ln_eps", c_float),
        ("act", c_int32),
        ("patch_bias", c_int32),
        ("pre_ln", c_int32),
        ("proj_dim", c_int32),
        ("dtype", c_int32),
        ("cta_group", c_int32),
    ]


KCLASSES = ("gemm", "attention", "layernorm", "other")


class KernelStats(ctypes.Structure):
    """Mirror of `vidil_kernel_stats`."""

    _fields_ = [("ms", ctypes.c_double * 4), ("flops", ctypes.c_double * 4), ("bytes", ctypes.c_double * 4),
                ("launches", c_int64 * 4)]


class TextCfg(ctypes.Structure):
    """Mirror of `vidil_text_cfg`."""

    _fields_ = [("vocab_size", c_int32), ("max_positions", c_int32), ("embed_dim", c_int32), ("depth", c_int32),
                ("num_heads", c_int32), ("mlp_dim", c_int32), ("ln_eps", c_float), ("act", c_int32), ("proj_dim", c_int32),
                ("dtype", c_int32), ("cta_group", c_int32)]


class MedCfg(ctypes.Structure):
    """Mirror of `vidil_med_cfg`."""

    _fields_ = [("vocab_size", c_int32), ("max_positions", c_int32), ("hidden", c_int32), ("depth", c_int32),
                ("num_heads", c_int32), ("mlp_dim", c_int32), ("encoder_width", c_int32), ("ln_eps", c_float),
                ("lm_head", c_int32), ("cls_out", c_int32), ("dtype", c_int32), ("cta_group", c_int32)]


# name -> (restype, argtypes); the single source of truth the symbol test checks against the header
SIGNATURES = {
    "vidil_abi_version": (c_int32, []),
    "vidil_last_error": (c_char_p, []),
    "vidil_kernel_launch_count": (c_int64, []),
    "vidil_encoder_create": (c_int32, [POINTER(EncoderCfg), POINTER(c_void_p)]),
    "vidil_encoder_destroy": (None, [c_void_p]),
    "vidil_encoder_load": (c_int32, [c_void_p, c_char_p, c_void_p, c_int64, c_void_p]),
    "vidil_encoder_check_loaded": (c_int32, [c_void_p]),
    "vidil_encoder_workspace_bytes": (c_size_t, [c_void_p, c_int32]),
    "vidil_encoder_tokens": (c_int32, [c_void_p]),
    "vidil_vit_forward": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vidil_vit_forward16": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vidil_clip_forward": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vidil_text_encoder_create": (c_int32, [POINTER(TextCfg), POINTER(c_void_p)]),
    "vidil_text_encoder_destroy": (None, [c_void_p]),
    "vidil_text_encoder_load": (c_int32, [c_void_p, c_char_p, c_void_p, c_int64, c_void_p]),
    "vidil_text_encoder_check_loaded": (c_int32, [c_void_p]),
    "vidil_text_encoder_workspace_bytes": (c_size_t, [c_void_p, c_int32, c_int32]),
    "vidil_clip_text_forward": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_size_t,
                                          c_void_p]),
    "vidil_med_create": (c_int32, [POINTER(MedCfg), POINTER(c_void_p)]),
    "vidil_med_destroy": (None, [c_void_p]),
    "vidil_med_load": (c_int32, [c_void_p, c_char_p, c_void_p, c_int64, c_void_p]),
    "vidil_med_check_loaded": (c_int32, [c_void_p]),
    "vidil_med_set_profiling": (c_int32, [c_void_p, c_int32]),
    "vidil_med_read_profile": (c_int32, [c_void_p, POINTER(KernelStats)]),
    "vidil_med_forward_workspace_bytes": (c_size_t, [c_void_p, c_int32, c_int32, c_int32, c_int32]),
    "vidil_med_forward": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                    c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vidil_med_generate_workspace_bytes": (c_size_t, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32]),
    "vidil_med_generate": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                     c_int32, c_int32, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vidil_med_sample": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                   c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vidil_op_sample": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                  c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vidil_op_beam_search_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32]),
    "vidil_op_beam_search": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_int32, c_int32, c_int32,
                                       c_int32, c_int32, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vidil_preprocess_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32, c_int32]),
    "vidil_preprocess_frames": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, POINTER(c_float), POINTER(c_float),
                                          c_void_p, c_void_p, c_size_t, c_void_p]),
    "vidil_clip_preprocess_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32, c_int32]),
    "vidil_clip_preprocess_frames": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, POINTER(c_float), POINTER(c_float),
                                               c_void_p, c_void_p, c_size_t, c_void_p]),
    "vidil_encoder_set_profiling": (c_int32, [c_void_p, c_int32]),
    "vidil_encoder_read_profile": (c_int32, [c_void_p, POINTER(KernelStats)]),
    "vidil_encoder_host_scratch_bytes": (c_size_t, [c_void_p, c_int32]),
    "vidil_vit_forward_host": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vidil_clip_forward_host": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vidil_encoder_host_pipeline_scratch_bytes": (c_size_t, [c_void_p, c_int32]),
    "vidil_encoder_host_submit": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_size_t, c_void_p]),
    "vidil_encoder_host_submit16": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_size_t, c_void_p]),
    "vidil_encoder_host_wait": (c_int32, [c_void_p, c_int32]),
    "vidil_sim_bank_create": (c_int32, [c_void_p, c_int32, c_int32, c_void_p, POINTER(c_void_p)]),
    "vidil_sim_bank_destroy": (None, [c_void_p]),
    "vidil_sim_bank_topk_workspace_bytes": (c_size_t, [c_void_p, c_int32]),
    "vidil_sim_bank_topk": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vidil_sim_topk_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32]),
    "vidil_sim_topk": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                 c_size_t, c_void_p]),
    "vidil_debug_set_attention_trace": (None, [c_void_p]),
    "vidil_op_linear_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32]),
    "vidil_op_linear": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                  c_int32, c_void_p, c_int32, c_void_p, c_size_t, c_void_p]),
    "vidil_op_layernorm": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_float, c_void_p]),
    "vidil_op_attention_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32]),
    "vidil_op_attention": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_float, c_int32, c_void_p,
                                     c_size_t, c_void_p]),
    "vidil_op_attention_causal": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_float, c_int32, c_void_p,
                                            c_size_t, c_void_p]),
}

_lib = None


def lib_path() -> str:
    return _build.LIB_PATH


def load() -> ctypes.CDLL:
    """Load (building first if needed) the native library; raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    _build.build()  # returns at once when build.stamp matches the sources; rebuilds a stale library (file-locked, atomic)
    lib = ctypes.CDLL(_build.LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means the .so is stale: rebuild with --force
        fn.restype = res
        fn.argtypes = args
    if lib.vidil_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libvidil_b200.so ABI {lib.vidil_abi_version()} != {ABI_VERSION}; run python -m vidil_b200.build --force")
    _lib = lib
    return lib


def last_error() -> str:
    return load().vidil_last_error().decode("utf-8", "replace")


def check(status: int, what: str) -> None:
    """Turn a non-zero status into the RuntimeError the reference's torch ops would have raised."""
    if status != 0:
        raise RuntimeError(f"{what}: {last_error()}")


def launch_count() -> int:
    return int(load().vidil_kernel_launch_count())
