"""ctypes binding of the C ABI declared in include/sodt_b200.h.

The shared library is built in-tree by ``build.py`` (nvcc, sm_100a).  Loading is lazy and
fails loudly: there is no fallback implementation of any entry point.
"""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SODT_B200_LIB") or os.path.join(_PKG, "libsodt_b200.so")     # override: kernel experiments only

_p = ctypes.c_void_p
_i = ctypes.c_int
_ll = ctypes.c_longlong
_f = ctypes.c_float
_d = ctypes.c_double
_sz = ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol of include/sodt_b200.h
SIGNATURES = {
    "sodt_version": (_i, []),
    "sodt_status_string": (ctypes.c_char_p, [_i]),
    "sodt_last_cuda_error": (ctypes.c_char_p, []),
    "sodt_built_for_sm": (_i, []),
    "sodt_window_attn_workspace_bytes": (_sz, [_i, _i, _i]),
    "sodt_window_attn_fwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _f, _f, _p, _sz, _p]),
    "sodt_window_attn_ex_fwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _f, _f, _p, _i, _p, _i, _p, _p, _p]),
    "sodt_window_attn_kernel_class": (_i, [_i, _i, _i, _i, _i, _i, _i, _i]),
    "sodt_window_attn_prepare": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _sz, _p]),
    "sodt_window_attn_fwd_prepared": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _f, _f, _p, _sz, _p]),
    "sodt_add_layernorm_fwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _ll, _i, _f, _i, _p]),
    "sodt_cattn_block_fwd": (_i, [_p, _p, _p, _p, _ll, _ll, _ll, _ll, _p, _p, _p,
                                  _i, _i, _i, _i, _i, _i, _i, _f, _f, _i, _p]),
    "sodt_frontend_fwd": (_i, [_p, _ll, _ll, _ll, _ll, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _i, _p]),
    "sodt_frontend_u8_fwd": (_i, [_p, _ll, _ll, _ll, _ll, _p, _ll, _ll, _ll, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _i, _p]),
    "sodt_frontend_embed_u8_supported": (_i, [_i, _i, _i, _i, _i, _i]),
    "sodt_frontend_embed_u8_fwd": (_i, [_p, _ll, _ll, _ll, _p, _ll, _ll, _p, _p, _p, _p, _p, _p, _p, _i, _p, _p,
                                        _i, _i, _i, _i, _i, _i, _f, _p]),
    "sodt_detect_decode": (_i, [_p, _ll, _ll, _ll, _ll, _p, _p, _p, _i, _i, _i, _i, _i, _f, _ll, _ll, _i, _p]),
    "sodt_nms_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "sodt_nms": (_i, [_p, _p, _i, _p, _p, _p, _p, _sz, _i, _i, _i, _f, _d, _i, _i, _i, _i, _i, _i, _f, _p]),
    "sodt_linear_supported": (_i, [_i, _i, _i, _i]),
    "sodt_linear_fwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "sodt_linear_strided_fwd": (_i, [_p, _i, _p, _i, _i, _p, _p, _p, _i, _i, _p, _i, _i, _i, _i, _i, _i, _p]),
    "sodt_linear_ln_fwd": (_i, [_p, _i, _p, _i, _f, _p, _p, _p, _p, _i, _i, _p, _i, _p, _i, _i, _i, _i, _i, _p]),
    "sodt_attn_block_supported": (_i, [_i, _i, _i, _i, _i, _i, _i, _i]),
    "sodt_attn_block_fwd": (_i, [_p, _p, _i, _f, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _f, _f, _p, _sz, _p]),
    "sodt_mlp_supported": (_i, [_i, _i, _i, _i]),
    "sodt_mlp_ln_fwd": (_i, [_p, _i, _p, _i, _f, _p, _p, _p, _p, _p, _p, _i, _p, _i, _i, _i, _i, _i, _p]),
    "sodt_row_stats": (_i, [_p, _ll, _p, _ll, _i, _f, _i, _p]),
    "sodt_stats_finalize": (_i, [_p, _i, _p, _ll, _i, _f, _p]),
    "sodt_conv2d_nhwc_supported": (_i, [_i, _i, _i, _i, _i, _i, _i, _i]),
    "sodt_conv2d_nhwc_fwd": (_i, [_p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "sodt_upcat_conv1x1_supported": (_i, [_i, _i, _i, _i, _i, _i, _i]),
    "sodt_upcat_conv1x1_fwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "sodt_patch_merge_linear_supported": (_i, [_i, _i, _i, _i, _i, _i]),
    "sodt_patch_merge_linear_fwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "sodt_bias_act_crop_nhwc": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "sodt_upsample2x_concat_nhwc": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "sodt_launch_count": (_ll, []),
    "sodt_reset_launch_count": (None, []),
}

_lib = None


class SodtError(RuntimeError):
    pass


def lib():
    """Returns the loaded library; raises SodtError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SodtError(
                f"{LIB_PATH} is missing: build it with `python small-object-detection-transformers_b200/build.py` "
                "(or __graft_entry__.build()).  There is no CPU fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(status, what):
    if status != 0:
        L = lib()
        msg = L.sodt_status_string(status).decode()
        if status == -4:
            msg += ": " + L.sodt_last_cuda_error().decode()
        raise SodtError(f"{what} failed: {msg} (status {status})")
