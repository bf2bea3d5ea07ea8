"""ctypes binding of libkeep_b200.so (include/keep_b200.h).

The library is the product path: there is no Python/CPU fallback.  `lib()` builds it on first use if the
toolkit is present and raises (loudly) otherwise; every wrapper raises `KeepB200Error` on a non-zero status.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

from . import build as _build


class KeepB200Error(RuntimeError):
    """A keepb200_* call returned a negative status; the message comes from keepb200_last_error()."""


class KeepB200Config(C.Structure):
    """Mirror of `struct KeepB200Config` (include/keep_b200.h)."""

    _fields_ = [
        ("struct_size", C.c_int32),
        ("img_size", C.c_int32),
        ("patch_size", C.c_int32),
        ("vit_width", C.c_int32),
        ("vit_depth", C.c_int32),
        ("vit_heads", C.c_int32),
        ("vit_mlp", C.c_int32),
        ("vit_ln_eps", C.c_float),
        ("proj_dim", C.c_int32),
        ("vocab_size", C.c_int32),
        ("hidden", C.c_int32),
        ("layers", C.c_int32),
        ("heads", C.c_int32),
        ("intermediate", C.c_int32),
        ("max_pos", C.c_int32),
        ("type_vocab", C.c_int32),
        ("bert_ln_eps", C.c_float),
        ("operand_dtype", C.c_int32),
    ]


_p = C.c_void_p
_i64 = C.c_int64
_int = C.c_int
_f = C.c_float
_sz = C.c_size_t

# name -> (restype, argtypes); exactly the declarations of include/keep_b200.h
SIGNATURES = {
    "keepb200_version": (_int, []),
    "keepb200_last_error": (C.c_char_p, []),
    "keepb200_create": (_int, [C.POINTER(KeepB200Config), _int, C.POINTER(_p)]),
    "keepb200_destroy": (None, [_p]),
    "keepb200_load_weight": (_int, [_p, C.c_char_p, _p, C.POINTER(_i64), _int, _p]),
    "keepb200_num_weights": (_int, [_p]),
    "keepb200_weight_name": (C.c_char_p, [_p, _int]),
    "keepb200_finalize": (_int, [_p]),
    "keepb200_workspace_bytes": (_sz, [_p, _int, _i64, _i64]),
    "keepb200_encode_image": (_int, [_p, _p, _int, _i64, _int, _p, _p, _sz, _p]),
    "keepb200_image_precision_is_high": (_int, [_int, _i64]),
    "keepb200_workspace_bytes_hw": (_sz, [_p, _i64, _i64, _i64, _int]),
    "keepb200_encode_image_hw": (_int, [_p, _p, _int, _i64, _i64, _i64, _int, _p, _p, _sz, _p]),
    "keepb200_preprocess_workspace_bytes": (_sz, [_i64, _i64, _i64, _int]),
    "keepb200_preprocess_u8": (_int, [_p, _i64, _i64, _i64, _int, _p, _p, _sz, _p]),
    "keepb200_encode_text": (_int, [_p, _p, _p, _p, _i64, _i64, _i64, _int, _p, _p, _sz, _p]),
    "keepb200_similarity": (_int, [_p, _i64, _i64, _p, _i64, _int, _f, _p, _p, _p, _sz, _p]),
    "keepb200_similarity_workspace_bytes": (_sz, [_i64, _i64]),
    "keepb200_prompt_scores_workspace_bytes": (_sz, [_i64, _i64, _i64, _i64]),
    "keepb200_prompt_scores": (_int, [_p, _i64, _i64, _p, _i64, _i64, _int, _p, _p, _sz, _p]),
    "keepb200_refine": (_int, [_p, _p, _i64, _i64, _i64, _int, _p, _p, _p, _sz, _p]),
    "keepb200_refine_workspace_bytes": (_sz, [_i64]),
    "keepb200_launch_count": (_i64, []),
    "keepb200_profile_begin": (_int, []),
    "keepb200_profile_end": (_int, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_i64), C.POINTER(_i64)]),
    "keepb200_profile_table": (C.c_char_p, []),
    "keepb200_op_gemm": (_int, [_p, _i64, _p, _i64, _int, _int, _int, _int, _int, _p, _p, _p, _i64, _p, _i64, _p, _int, _p]),
    "keepb200_op_gemm_resid_stats": (_int, [_p, _p, _int, _int, _int, _int, _p, _p, _p, _p, _p, _p]),
    "keepb200_op_gemm_ln": (_int, [_p, _p, _int, _int, _int, _int, _int, _p, _p, _p, _f, _p, _p]),
    "keepb200_op_fold_ln": (_int, [_p, _int, _int, _p, _p, _p, _p, _int, _p, _p, _p]),
    "keepb200_op_pos_resample": (_int, [_p, _int, _int, _int, _int, _p, _p]),
    "keepb200_op_layernorm": (_int, [_p, _i64, _i64, _int, _p, _p, _f, _p, _int, _p, _i64, _i64, _p]),
    "keepb200_op_attention": (_int, [_p, _p, _int, _int, _int, _int, _p, _i64, _f, _i64, _i64, _p]),
    "keepb200_op_gemm_split": (_int, [_p, _i64, _p, _i64, _int, _int, _int, _int, _int, _int, _p, _p, _i64, _p, _i64, _i64, _p]),
    "keepb200_op_cast_hilo": (_int, [_p, _p, _i64, _int, _int, _p]),
    "keepb200_op_visual_head": (_int, [_p, _i64, _i64, _int, _p, _p, _f, _p, _p, _int, _p, _p, _int, _p, _p]),
    "keepb200_op_pooler": (_int, [_p, _i64, _i64, _int, _p, _p, _p, _p]),
    "keepb200_debug_set_ln_fuse": (_int, [_p, _int]),
    "keepb200_debug_layer_dump": (_int, [_p, _p, _sz]),
    "keepb200_debug_attention_trace": (_int, [_p]),
    "keepb200_op_act_l2norm": (_int, [_p, _i64, _int, _int, _p, _p]),
}

ABI_VERSION = 2  # KEEPB200_ABI_VERSION of include/keep_b200.h
PRECISION = {"auto": 0, "high": 1, "fast": 2, "balanced": 3}  # KEEPB200_PRECISION_*

_LIB = None


def lib_path() -> Path:
    return _build.LIB_PATH


def lib() -> C.CDLL:
    """Load (building if stale and nvcc is available) libkeep_b200.so. Raises if it cannot be had."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB_PATH
    override = os.environ.get("KEEPB200_LIB")  # A/B measurements against another build of the same ABI
    if override:
        path = Path(override)
    else:
        # build() takes an inter-process lock around the stale check, the compile and the atomic replace of the
        # library, so the ranks of a torchrun job never load a half-written file
        try:
            path = _build.build()
        except _build.ToolchainMissing as e:  # no nvcc on this box: only a complete prebuilt in-tree library will do
            if not path.exists():
                raise KeepB200Error(
                    f"libkeep_b200.so is missing and could not be built ({e}); "
                    "keep_b200 has no CPU or PyTorch fallback"
                ) from e
        except Exception as e:  # a failed compile must never fall back to a stale library
            raise KeepB200Error(f"building libkeep_b200.so failed: {e}") from e
    handle = C.CDLL(str(path))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if handle.keepb200_version() != ABI_VERSION:
        raise KeepB200Error(f"ABI version mismatch: library {handle.keepb200_version()} != binding {ABI_VERSION}")
    _LIB = handle
    return handle


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = lib().keepb200_last_error()
        raise KeepB200Error(f"{what or 'keepb200 call'} failed ({status}): {msg.decode() if msg else '?'}")


def ptr(t) -> int | None:
    """Device/host pointer of a torch tensor (None passes NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr(device=None) -> int:
    import torch

    return torch.cuda.current_stream(device).cuda_stream
