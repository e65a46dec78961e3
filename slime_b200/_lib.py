"""ctypes binding of libslime_b200.so (include/slime_b200.h).

The library is the product; there is no Python/PyTorch fallback.  Importing this module on a box
where the library is missing raises immediately, and every wrapper raises RuntimeError with the
library's own message when an entry point returns a negative code.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libslime_b200.so"
# element type -> build of the library (same sources, -DSLIME_FP16 for half; slime_b200/build.py)
LIB_PATHS = {"bf16": LIB_PATH, "fp16": _PKG / "libslime_b200_fp16.so"}
ELEM_DTYPE_CODE = {"bf16": 0, "fp16": 2}


def variant_of(dtype) -> str:
    """'bf16' / 'fp16' from a torch dtype or a string (None = bf16)."""
    if dtype is None:
        return "bf16"
    name = str(dtype).replace("torch.", "")
    if name in ("bf16", "bfloat16"):
        return "bf16"
    if name in ("fp16", "float16", "half"):
        return "fp16"
    raise ValueError(f"slime_b200 computes in bfloat16 or float16, not {dtype}")


def torch_dtype(variant: str):
    import torch

    return torch.float16 if variant == "fp16" else torch.bfloat16


SLIME_FLAG_LEFT_PAD = 1
SLIME_FLAG_USE_GLOBAL_ONLY = 2
SLIME_FLAG_USE_LOCAL_ONLY = 4
SLIME_FLAG_ROPE_INTERLEAVED = 8
SLIME_FLAG_ROUTER_QFORMER = 16
SLIME_FLAG_NORM_FOLDED = 32

EPI_NONE, EPI_QUICK_GELU, EPI_GELU_ERF, EPI_SWIGLU, EPI_ROPE = 0, 1, 2, 3, 4


class ModelDesc(C.Structure):
    """Mirror of `slime_model_desc` (include/slime_b200.h)."""

    _fields_ = [
        ("vit_hidden", C.c_int32),
        ("vit_layers_used", C.c_int32),
        ("vit_heads", C.c_int32),
        ("vit_mlp", C.c_int32),
        ("vit_image", C.c_int32),
        ("vit_patch", C.c_int32),
        ("vit_ln_eps", C.c_float),
        ("rs_local_queries", C.c_int32),
        ("rs_global_queries", C.c_int32),
        ("rs_ln_eps", C.c_float),
        ("mm_learnable_gated", C.c_int32),
        ("hidden", C.c_int32),
        ("layers", C.c_int32),
        ("heads", C.c_int32),
        ("kv_heads", C.c_int32),
        ("head_dim", C.c_int32),
        ("mlp", C.c_int32),
        ("vocab", C.c_int32),
        ("rope_theta", C.c_float),
        ("rms_eps", C.c_float),
        ("max_pos", C.c_int32),
        ("top_p", C.c_float),
        ("temp", C.c_float),
        ("image_token", C.c_int64),
        ("sep_token", C.c_int64),
        ("max_len", C.c_int32),
        ("flags", C.c_uint32),
    ]


class ResizeJob(C.Structure):
    """Mirror of `slime_resize_job` (include/slime_b200.h)."""

    _fields_ = [
        ("src_offset", C.c_int64),
        ("src_w", C.c_int32), ("src_h", C.c_int32),
        ("virt_w", C.c_int32), ("virt_h", C.c_int32), ("virt_x", C.c_int32), ("virt_y", C.c_int32),
        ("out_w", C.c_int32), ("out_h", C.c_int32),
        ("canvas_w", C.c_int32), ("canvas_h", C.c_int32),
        ("paste_x", C.c_int32), ("paste_y", C.c_int32),
        ("first_crop", C.c_int32),
        ("fill", C.c_uint8 * 4),
    ]


_vp, _i, _i64, _f, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t

# name -> (restype, argtypes); every symbol declared in include/slime_b200.h
SIGNATURES = {
    "slime_version": (_i, []),
    "slime_elem_dtype": (_i, []),
    "slime_last_error": (C.c_char_p, []),
    "slime_ctx_create": (_i, [C.POINTER(_vp), _i, C.POINTER(ModelDesc)]),
    "slime_ctx_destroy": (None, [_vp]),
    "slime_ctx_set_weight": (_i, [_vp, C.c_char_p, _vp, _i64, _i64]),
    "slime_finalize_workspace_bytes": (_sz, [_vp]),
    "slime_ctx_finalize_weights": (_i, [_vp, _vp, _sz, _vp]),
    "slime_vision_tower_workspace_bytes": (_sz, [_vp, _i]),
    "slime_vision_tower_fwd": (_i, [_vp, _vp, _i, _vp, _vp, _sz, _vp]),
    "slime_vision_tower_fwd_split": (_i, [_vp, _vp, _i, _i, _vp, _vp, _sz, _vp]),
    "slime_resampler_workspace_bytes": (_sz, [_vp, _i, _i]),
    "slime_resampler_fwd": (_i, [_vp, _i, _vp, _i, _vp, _vp, _sz, _vp]),
    "slime_projector_workspace_bytes": (_sz, [_vp, _i]),
    "slime_projector_fwd": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "slime_gated_projector_workspace_bytes": (_sz, [_vp, _i]),
    "slime_gated_projector_fwd": (_i, [_vp, _vp, _i, _vp, _vp, _sz, _vp]),
    "slime_router_workspace_bytes": (_sz, [_vp, _i, _i, _i]),
    "slime_router_fwd": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "slime_router_fwd_embeds": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "slime_router_select": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "slime_splice_plan_ints": (_sz, [_i, _i]),
    "slime_splice_plan": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "slime_splice_plan_async": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "slime_splice_check": (_i, [_vp, _vp]),
    "slime_splice_gather": (_i, [_vp, _vp, _i, _i, _vp, _vp, _i, _i64, _vp, _i64, _vp, _i, _i, _vp, _vp, _i, _vp]),
    "slime_splice_pad": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "slime_decoder_workspace_bytes": (_sz, [_vp, _i, _i]),
    "slime_decoder_prefill_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "slime_kv_cache_bytes": (_sz, [_vp, _i, _i]),
    "slime_decoder_set_kv_cache": (_i, [_vp, _vp, _i, _i]),
    "slime_decoder_decode_workspace_bytes": (_sz, [_vp, _i]),
    "slime_decoder_decode_fwd": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _sz, _vp]),
    "slime_preprocess_workspace_bytes": (_sz, [_vp, _i]),
    "slime_preprocess_fwd": (_i, [_vp, _vp, _i, _i, _vp, _vp, _i, _vp, _sz, _vp]),
    "slime_op_gemm": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp, _i, _vp, _vp, _i, _vp]),
    "slime_op_gemm_skinny": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _i, _i, _vp, _sz, _vp, _vp, _f,
                                  _vp, _vp, _i, _i, _i, _vp]),
    "slime_op_gemm_skinny_fused": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _i, _i, _vp, _sz, _vp, _vp,
                                        _f, _vp, _vp, _i, _i, _i, _vp, _vp, _f, _vp]),
    "slime_op_decode_attention": (_i, [_vp, _i, _vp, _vp, _i, _vp, _i, _i, _i, _i, _f, _vp, _i, _i, _vp, _vp]),
    "slime_op_decode_attention_fused": (_i, [_vp, _i, _vp, _vp, _i, _vp, _i, _i, _i, _i, _f, _vp, _i, _i, _vp, _vp, _vp]),
    "slime_op_attention": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _i64, _i64, _i64, _i, _i, _i,
                                _i, _f, _i, _i64, _i64, _i, _vp]),
    "slime_op_layernorm": (_i, [_vp, _vp, _vp, _vp, _i, _i, _f, _vp]),
    "slime_op_rmsnorm": (_i, [_vp, _vp, _vp, _i, _i, _f, _vp]),
    "slime_op_rope": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "slime_op_qkv_rope": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "slime_gemm_set_2cta_mode": (_i, [_i]),
    "slime_gemm_set_epi_mode": (_i, [_i]),
    "slime_gemm_set_skinny_mode": (_i, [_i]),
    "slime_decode_attention_set_mode": (_i, [_i]),
    "slime_set_pdl_mode": (_i, [_i]),
    "slime_set_decode_prefetch": (_i, [_i]),
    "slime_gemm_set_tile_n": (_i, [_i]),
    "slime_set_prefill_pdl": (_i, [_i]),
    "slime_set_decode_fused": (_i, [_i]),
    "slime_attention_set_trace": (_i, [_vp]),
    "slime_attention_set_poly": (_i, [_i]),
    "slime_comm_unique_id": (_i, [_vp]),
    "slime_comm_init": (_i, [C.POINTER(_vp), _vp, _i, _i]),
    "slime_comm_nccl_version": (_i, []),
    "slime_allgather_logits": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "slime_comm_destroy": (None, [_vp]),
    "slime_launch_count": (C.c_longlong, []),
    "slime_profile_enable": (_i, [_i]),
    "slime_profile_collect": (_i, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
}

_libs = {}


def load(dtype=None) -> C.CDLL:
    """dlopen the build of the library for `dtype` (default bfloat16) and bind every declared symbol
    (raises if anything is missing)."""
    variant = variant_of(dtype)
    if variant in _libs:
        return _libs[variant]
    path = LIB_PATHS[variant]
    if not path.exists():
        raise RuntimeError(
            f"{path} is missing: build it with `python -m slime_b200.build` "
            "(slime_b200 has no PyTorch/CPU fallback path)")
    lib = C.CDLL(str(path))  # RTLD_LOCAL: the two builds export the same names and must not see each other
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.slime_version() != 1:
        raise RuntimeError(f"{path.name} ABI version {lib.slime_version()} != 1")
    if lib.slime_elem_dtype() != ELEM_DTYPE_CODE[variant]:
        raise RuntimeError(f"{path.name} was built for element type code {lib.slime_elem_dtype()}")
    _libs[variant] = lib
    return lib


def last_error(lib=None) -> str:
    """The message of the last failing call on this thread.  Each build keeps its own; without `lib` the
    non-empty one is returned."""
    libs = [lib] if lib is not None else list(_libs.values()) or [load()]
    msgs = [l.slime_last_error().decode("utf-8", "replace") for l in libs]
    return next((m for m in msgs if m), "")


def check(rc: int, what: str = "", lib=None) -> None:
    if rc != 0:
        raise RuntimeError(f"slime_b200 {what} failed (code {rc}): {last_error(lib)}")


def ptr(t) -> C.c_void_p:
    """Raw device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def stream_ptr() -> C.c_void_p:
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
