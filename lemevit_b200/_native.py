"""ctypes binding of liblemevit_b200.so (C ABI in include/lemevit_b200.h).

The library is built in-tree by ``lemevit_b200.build`` (nvcc, sm_100a).  There is no Python or CPU
fallback: if the library is missing and cannot be built, importing a symbol from here raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import build as _build

LMV_OK = 0
LMV_ERR_INVALID, LMV_ERR_UNSUPPORTED, LMV_ERR_CUDA, LMV_ERR_OOM = -1, -2, -3, -4
DTYPE_BF16, DTYPE_F32, DTYPE_U8, DTYPE_U8_NHWC = 0, 1, 2, 3
MAX_STAGES = 8


class Config(C.Structure):
    _fields_ = [
        ("num_stages", C.c_int),
        ("depth", C.c_int * MAX_STAGES),
        ("embed_dim", C.c_int * MAX_STAGES),
        ("mlp_hidden", C.c_int * MAX_STAGES),
        ("attn_type", C.c_char * MAX_STAGES),
        ("head_dim", C.c_int),
        ("queries_len", C.c_int),
        ("num_classes", C.c_int),
        ("in_chans", C.c_int),
        ("backbone", C.c_int),
    ]


class ProfileEntry(C.Structure):
    _fields_ = [("name", C.c_char_p), ("launches", C.c_longlong), ("device_ms", C.c_double), ("flops", C.c_double),
                ("bytes", C.c_double)]


class Tensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("numel", C.c_int64), ("dtype", C.c_int)]


_vp, _i, _f, _ll, _sz = C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_size_t

# name -> (restype, argtypes); every symbol declared in include/lemevit_b200.h
SIGNATURES = {
    "lmv_last_error": (C.c_char_p, []),
    "lmv_version": (_i, []),
    "lmv_packed_tensor_count": (_i, [C.POINTER(Config)]),
    "lmv_plan_create": (_i, [C.POINTER(Config), C.POINTER(Tensor), _i, C.POINTER(_vp)]),
    "lmv_plan_destroy": (None, [_vp]),
    "lmv_plan_set_chunk": (_i, [_vp, _i]),
    "lmv_plan_set_debug_simt": (_i, [_vp, _i]),
    "lmv_conv3x3s2": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "lmv_plan_set_input_norm": (_i, [_vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "lmv_plan_set_tap": (_i, [_vp, _i, _i, _vp, _vp]),
    "lmv_plan_set_option": (_i, [_vp, C.c_char_p, _i]),
    "lmv_plan_set_profile": (_i, [_vp, _i]),
    "lmv_plan_get_profile": (_i, [_vp, _vp, _i]),
    "lmv_plan_profile_report": (_i, [_vp, C.c_char_p, _i]),
    "lmv_workspace_bytes": (_sz, [_vp, _i, _i, _i]),
    "lmv_launch_count": (_i, [_vp, _i, _i, _i]),
    "lmv_forward_cls": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp, _i, _vp]),
    "lmv_forward_cls_features": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _sz, _vp, _vp, _i, _vp]),
    "lmv_forward_features": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _sz, C.POINTER(_vp), _i, _i, _vp]),
    "lmv_linear": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "lmv_linear_fused": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _f, _vp, _i, _vp]),
    "lmv_linear_stats_parts": (_i, [_i, _i]),
    "lmv_mlp_fused": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _f, _i, _i, _i, _vp]),
    "lmv_linear_simt": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "lmv_posembed_layernorm": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp]),
    "lmv_layernorm": (_i, [_vp, _vp, _vp, _vp, _i, _i, _f, _i, _i, _i, _i, _vp]),
    "lmv_attention": (_i, [_vp, _ll, _i, _vp, _ll, _i, _vp, _ll, _i, _vp, _ll, _i, _i, _i, _i, _i, _f, _i, _vp]),
    "lmv_attention_self_workspace": (_sz, [_i, _i, _i]),
    "lmv_attention_self": (_i, [_vp, _ll, _i, _vp, _ll, _i, _vp, _ll, _i, _vp, _ll, _i, _i, _i, _i, _i, _f, _vp, _sz, _vp]),
    "lmv_attention_meta_workspace": (_sz, [_i, _i, _i, _i]),
    "lmv_attention_meta": (_i, [_vp, _ll, _i, _vp, _ll, _i, _vp, _ll, _i, _vp, _ll, _i, _i, _i, _i, _i, _f, _vp, _sz, _vp]),
    "lmv_dca_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "lmv_dca_block": (_i, [_i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _f,
                           _vp, _sz, _i, _vp]),
    "lmv_stem_im2col": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp]),
    "lmv_stem_conv1": (_i, [_vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "lmv_im2col_3x3s2": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "lmv_tail": (_i, [_vp, _ll, _i, _vp, _ll, _i, _i, _vp, _vp, _vp, _vp, _f, _vp, _i, _vp]),
    "lmv_tokens_to_nchw": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
}

_lock = threading.Lock()
_lib = None


def library_path() -> str:
    return _build.LIB


def load() -> C.CDLL:
    """Load (building first if the sources are newer) the native library.  Raises on failure."""
    global _lib
    with _lock:
        if _lib is None:
            path = _build.LIB
            override = os.environ.get("LEMEVIT_B200_LIB")   # debug builds (e.g. the GEMM cycle-trace build) by explicit path
            if override:
                if not os.path.isfile(override):
                    raise RuntimeError(f"LEMEVIT_B200_LIB={override} does not exist")
                path = override
            elif not os.path.isfile(path) or (os.environ.get("LEMEVIT_B200_REBUILD") == "1"):
                path = _build.build(force=True)
            else:
                try:
                    path = _build.build(force=False)   # rebuild only when stale and nvcc is present
                except RuntimeError:
                    pass
            lib = C.CDLL(path)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)       # AttributeError if the .so does not export the symbol
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


class NativeError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"lemevit_b200 native error {code}: {message}")
        self.code = code


def check(rc: int) -> None:
    """Raise RuntimeError for a non-zero status (OOM keeps the text timm's batch-size retry looks for:
    reference benchmark.py:734-745)."""
    if rc != LMV_OK:
        msg = load().lmv_last_error().decode("utf-8", "replace")
        raise NativeError(rc, msg)


def make_config(depth, embed_dim, mlp_hidden, attn_type, head_dim, queries_len, num_classes, in_chans, backbone) -> Config:
    cfg = Config()
    n = len(attn_type)
    if n > MAX_STAGES:
        raise ValueError("too many stages")
    cfg.num_stages = n
    for i in range(n):
        cfg.depth[i] = int(depth[i])
        cfg.embed_dim[i] = int(embed_dim[i])
        cfg.mlp_hidden[i] = int(mlp_hidden[i])
    cfg.attn_type = "".join(attn_type).encode()
    cfg.head_dim, cfg.queries_len = int(head_dim), int(queries_len)
    cfg.num_classes, cfg.in_chans, cfg.backbone = int(num_classes), int(in_chans), int(bool(backbone))
    return cfg
