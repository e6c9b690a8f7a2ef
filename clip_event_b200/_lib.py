"""ctypes binding of libclip_event_b200.so (the C ABI in include/clip_event_b200.h).

There is no CPU fallback: if the library is missing and cannot be built, or a call fails,
a RuntimeError is raised (the reference's own error style, model_clip.py:631, engine.py:149).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CE_LIB_PATH") or os.path.join(_PKG, "libclip_event_b200.so")

CE_F32, CE_BF16 = 0, 1
CE_MASK_NUM_I64, CE_MASK_PAD_U8 = 0, 1
CE_IMG_CE_OVERBATCH, CE_IMG_CE_INSTANCE, CE_IMG_BCE_INSTANCE = 0, 1, 2

_vp, _i, _i64, _f, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t

# name -> (restype, argtypes); mirrors include/clip_event_b200.h one to one
SIGNATURES = {
    "ce_version": (_i, []),
    "ce_last_error": (C.c_char_p, []),
    "ce_device_check": (_i, []),
    "ce_contrastive_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "ce_contrastive_fwd_partial": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i64, _i, _i, _i64, _i, _vp, _vp, _vp, _sz, _vp]),
    "ce_contrastive_fwd_finish": (_i, [_vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "ce_contrastive_bwd_partial": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i64, _i, _i, _i64, _i, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ce_contrastive_bwd_finish": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "ce_contrastive_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "ce_contrastive_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ce_similarity_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "ce_similarity_logits": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "ce_ot_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "ce_ot_fwd_bwd": (_i, [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i, _i, _i, _i, _i, _i, _f, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ce_ot_cost_matrix": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp]),
    "ce_ot_ipot": (_i, [_vp, _vp, _vp, _i, _i, _i, _f, _i, _i, _vp, _vp]),
    "ce_ot_trace": (_i, [_vp, _i, _i, _vp, _vp]),
    "ce_scale_inplace": (_i, [_vp, _i64, _i64, _i64, _i, _vp, _vp]),
    "ce_scale_inplace_same": (_i, [_vp, _i64, _i, _vp, _vp, _vp]),
    "ce_head_losses_cast": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _i, _vp]),
    "ce_head_step_scale": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _vp, _i, _vp]),
    "ce_dense_ce_workspace_bytes": (_sz, [_i]),
    "ce_dense_ce_fwd": (_i, [_vp, _i64, _i64, _i, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "ce_dense_ce_bwd": (_i, [_vp, _i64, _i64, _i, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _i64, _vp, _vp]),
    "ce_ot_fwd_bwd_packed": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _i, _i, _f, _vp, _vp, _vp, _vp, _vp]),
    "ce_proj_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "ce_proj_fwd": (_i, [_vp, _i64, _vp, _vp, _vp, _f, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "ce_proj_bwd": (_i, [_vp, _i64, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ce_p2p_gather": (_i, [_vp, _i, _i64, _vp, _vp]),
    "ce_p2p_reduce_f32": (_i, [_vp, _i, _i64, _i64, _vp, _i64, _i, _vp, _vp]),
    "ce_head_param_step": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _f, _i, _f, _f, _f, _f, _f, _vp, _vp]),
    "ce_debug_launch_count": (C.c_ulonglong, []),
    "ce_debug_gemm": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "ce_debug_gemm_pair": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
}

_lock = threading.Lock()
_lib = None


def load(build_if_missing: bool = True):
    """Load (building first if the .so is absent and nvcc exists) and return the ctypes handle."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            if not build_if_missing:
                raise RuntimeError("clip_event_b200: %s is missing (run __graft_entry__.build())" % LIB_PATH)
            from . import build as _build
            _build.build()
        try:
            lib = C.CDLL(LIB_PATH)
        except OSError as e:  # pragma: no cover
            raise RuntimeError("clip_event_b200: cannot load %s: %s (no CPU fallback exists)" % (LIB_PATH, e))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    return load().ce_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise RuntimeError("clip_event_b200 %s failed (code %d): %s" % (what, rc, last_error()))


def dtype_code(t) -> int:
    import torch
    if t == torch.float32:
        return CE_F32
    if t == torch.bfloat16:
        return CE_BF16
    raise RuntimeError("clip_event_b200 supports float32 and bfloat16 embeddings, got %s" % (t,))


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("clip_event_b200 runs on CUDA (sm_100a) tensors only; got a %s tensor. "
                               "There is no CPU path." % t.device)


def ptr(t):
    return 0 if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
