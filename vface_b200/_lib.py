"""ctypes binding of libvface_b200.so (the C-ABI declared in include/vface_b200.h).

There is no fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvface_b200.so")
ABI_VERSION = 10

VF_F32 = 0
VF_BF16 = 1

_lib = None
_lock = threading.Lock()

_c = ctypes
_vp, _ll, _i, _f, _dbl = _c.c_void_p, _c.c_longlong, _c.c_int, _c.c_float, _c.c_double

# name -> (restype, argtypes); mirrors include/vface_b200.h one to one.
SIGNATURES = {
    "vf_abi_version": (_i, []),
    "vf_last_error": (_c.c_char_p, []),
    "vf_attn_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _ll, _ll, _ll, _ll, _f,
                         _vp, _vp, _i, _ll, _ll, _i, _vp]),
    "vf_fsai_blend": (_i, [_vp, _vp, _vp, _ll, _i, _i, _ll, _ll, _ll, _i, _vp]),
    "vf_fsai_blend2": (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _ll, _ll, _ll, _ll, _ll, _i, _vp]),
    "vf_flow_warp_blend": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _ll, _ll, _ll, _dbl, _i, _vp, _vp]),
    "vf_ddim_cfg_step": (_i, [_vp, _vp, _vp, _vp, _vp, _f, _f, _f, _f, _f, _vp, _ll, _i, _vp]),
    "vf_ddim_invert_step": (_i, [_vp, _vp, _vp, _vp, _f, _f, _f, _ll, _i, _vp]),
    "vf_group_norm_workspace_floats": (_ll, [_i, _i, _i]),
    "vf_group_norm_nhwc": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _i, _vp]),
    "vf_group_norm_nhwc_cat": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _i, _i, _vp]),
    "vf_add_layer_norm": (_i, [_vp, _vp, _vp, _ll, _vp, _vp, _vp, _vp, _ll, _i, _f, _i, _vp]),
    "vf_geglu": (_i, [_vp, _vp, _ll, _i, _ll, _i, _vp]),
    "vf_linear_geglu": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _i, _ll, _i, _vp]),
    "vf_upsample_nearest2x_nhwc": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "vf_add_bias": (_i, [_vp, _vp, _vp, _ll, _vp, _ll, _i, _i, _vp]),
    "vf_conv3x3_out_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "vf_linear_residual": (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _ll, _ll, _ll, _vp, _ll, _i, _vp]),
    "vf_linear_proj_supported": (_i, [_ll, _i, _i]),
    "vf_linear_proj": (_i, [_vp, _vp, _vp, _ll, _vp, _vp, _f, _vp, _i, _vp, _vp, _ll, _i, _i, _ll, _ll, _ll, _i, _vp]),
    "vf_linear_residual_batched": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _ll, _i, _i, _ll, _ll, _ll, _vp, _ll, _i, _vp]),
}


def load():
    """Load the library once; raise loudly if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m vface_b200.build` "
                "(vface_b200 has no CPU or PyTorch fallback for its kernels)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)      # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        got = lib.vf_abi_version()
        if got != ABI_VERSION:
            raise RuntimeError(f"libvface_b200.so ABI {got} != expected {ABI_VERSION}; rebuild")
        _lib = lib
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().vf_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")
