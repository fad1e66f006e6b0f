"""ctypes binding of libmovedepth_b200.so (the C ABI declared in include/movedepth_b200.h).

There is no fallback: if the shared library is missing or a symbol of the header is not
exported, importing the ops raises.  Build with `python -m movedepth_b200.build`.
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "libmovedepth_b200.so")
HEADER = os.path.join(ROOT, "include", "movedepth_b200.h")

_P = ctypes.c_void_p
_I = ctypes.c_int
_F = ctypes.c_float
_LL = ctypes.c_longlong
_D = ctypes.c_double

# argument types per entry point, in header order
SIGNATURES = {
    "mvd_version": ([], _I),
    "mvd_last_error_string": ([], ctypes.c_char_p),
    "mvd_sm_count": ([], _I),
    "mvd_costvol_grouped_fwd": ([_P] * 9 + [_I] * 8 + [_P], _I),
    "mvd_costvol_grouped_bwd": ([_P] * 11 + [_I] * 8 + [_P], _I),
    "mvd_costvol_full_fwd": ([_P] * 7 + [_I] * 5 + [_P], _I),
    "mvd_costvol_full_bwd": ([_P] * 9 + [_I] * 5 + [_P], _I),
    "mvd_regress_fwd": ([_P] * 7 + [_I] * 4 + [_P], _I),
    "mvd_regress_bwd": ([_P] * 7 + [_I] * 4 + [_P], _I),
    "mvd_convex_up_fwd": ([_P] * 3 + [_I] * 4 + [_P], _I),
    "mvd_convex_up_bwd": ([_P] * 5 + [_I] * 4 + [_P], _I),
    "mvd_photometric_fwd": ([_P] * 8 + [_I] * 3 + [_F, _I, _P], _I),
    "mvd_photometric_bwd": ([_P] * 10 + [_I] * 3 + [_F, _P], _I),
    "mvd_reproj_select_fwd": ([_P] * 8 + [_LL, _P], _I),
    "mvd_reproj_select_bwd": ([_P] * 5 + [_LL, _P], _I),
    "mvd_disp_to_depth_fwd": ([_P, _P] + [_I] * 5 + [_F, _F, _P], _I),
    "mvd_disp_to_depth_bwd": ([_P, _P, _P] + [_I] * 5 + [_F, _P], _I),
    "mvd_smooth_loss_workspace_bytes": ([_I], _LL),
    "mvd_smooth_loss_fwd": ([_P] * 4 + [_I] * 4 + [_P], _I),
    "mvd_smooth_loss_bwd": ([_P] * 6 + [_I] * 4 + [_P], _I),
    "mvd_masked_smooth_l1_fwd": ([_P] * 6 + [_I] * 7 + [_F, _P], _I),
    "mvd_masked_smooth_l1_bwd": ([_P] * 7 + [_LL, _F, _P], _I),
    "mvd_conv2d_small_supported": ([_I] * 4, _I),
    "mvd_conv2d_small_fwd": ([_P] * 3 + [_I] * 8 + [_P], _I),
    "mvd_conv2d_small_dgrad": ([_P] * 3 + [_I] * 8 + [_P], _I),
    "mvd_conv2d_small_wgrad_workspace_bytes": ([_I] * 8, _LL),
    "mvd_conv2d_small_wgrad": ([_P] * 4 + [_LL] + [_I] * 8 + [_P], _I),
    "mvd_decoder_prep_fwd": ([_P] * 5 + [_I] * 7 + [_P], _I),
    "mvd_decoder_prep_bwd": ([_P] * 6 + [_I] * 7 + [_P], _I),
    "mvd_gather_chunk": ([], _I),
    "mvd_gather_segments": ([_P, _P, _I, _P, _P], _I),
    "mvd_maxpool3x3s2_fwd": ([_P] * 3 + [_I] * 4 + [_P], _I),
    "mvd_maxpool3x3s2_bwd": ([_P] * 3 + [_I] * 4 + [_P], _I),
    "mvd_pose_matrix_fwd": ([_P] * 3 + [_I] * 2 + [_P], _I),
    "mvd_pose_matrix_bwd": ([_P] * 5 + [_I] * 2 + [_P], _I),
    "mvd_resample_u8": ([_P] * 4 + [_I, _LL, _I, _I, _I, _P, _LL, _P], _I),
    "mvd_flip_copy_u8": ([_P, _P] + [_I] * 4 + [_P, _P], _I),
    "mvd_u8_to_tensor": ([_P, _P] + [_I] * 3 + [_P], _I),
    "mvd_jitter_blend_u8": ([_P] + [_I] * 4 + [_P, _P, _P, _P], _I),
    "mvd_jitter_hue_u8": ([_P] + [_I] * 3 + [_P, _P, _P], _I),
    "mvd_split_tf32": ([_P, _P, _LL, _I, _I, _P], _I),
    "mvd_conv3d_c16o1_fwd": ([_P] * 3 + [_I] * 4 + [_P], _I),
    "mvd_conv3d_c16o1_dgrad": ([_P] * 3 + [_I] * 4 + [_P], _I),
    "mvd_conv3d_c16o1_wgrad_workspace_bytes": ([_I] * 4, _LL),
    "mvd_conv3d_c16o1_wgrad": ([_P] * 4 + [_LL] + [_I] * 4 + [_P], _I),
    "mvd_conv3d_c16c16": ([_P] * 3 + [_I] * 6 + [_P], _I),
    "mvd_conv3d_c16c16_tc": ([_P] * 4 + [_I] * 7 + [_P], _I),
    "mvd_conv3d_c16c16_wgrad_tc_workspace_bytes": ([_I] * 4, _LL),
    "mvd_conv3d_c16c16_wgrad_tc": ([_P] * 4 + [_LL] + [_I] * 4 + [_P], _I),
    "mvd_conv3d_c16c16_wgrad_workspace_bytes": ([_I] * 4, _LL),
    "mvd_conv3d_c16c16_wgrad": ([_P] * 4 + [_LL] + [_I] * 4 + [_P], _I),
    "mvd_bn_stats": ([_P, _LL, _I, _P, _P, _I, _I, _I, _P], _I),
    "mvd_bn_finalize": ([_P, _D, _P, _P, _P, _P, _F, _F, _P, _I, _P, _P], _I),
    "mvd_bn_apply": ([_P, _P, _P, _P, _LL, _I, _I, _P], _I),
    "mvd_bn_bwd_reduce": ([_P, _P, _P, _P, _P, _P, _LL, _I, _I, _P, _I, _I, _I, _P], _I),
    "mvd_bn_bwd_apply": ([_P] * 6 + [_D] + [_P] * 4 + [_LL, _I, _I, _P], _I),
    "mvd_bn_workspace_doubles": ([], _I),
    "mvd_bn_fwd_fused": ([_P] * 7 + [_F, _F, _D, _P, _P, _LL, _I, _I, _P, _P, _I, _I, _I, _P], _I),
    "mvd_bn_bwd_fused": ([_P] * 5 + [_D] + [_P] * 5 + [_LL, _I, _I, _P, _P, _I, _I, _I, _P], _I),
    "mvd_peer_allreduce_buffer_bytes": ([_I, _I], _LL),
    "mvd_peer_allreduce_f64": ([_P, _P, _I, _P, _I, _I, _I, _P], _I),
    "mvd_event_create": ([], _P),
    "mvd_event_record": ([_P, _P, _I], _I),
    "mvd_event_elapsed_ms": ([_P, _P, ctypes.POINTER(ctypes.c_float)], _I),
    "mvd_event_destroy": ([_P], _I),
    "mvd_adam_step": ([_P] * 4 + [_LL] + [_F] * 6 + [_P], _I),
}


def declared_symbols(header=HEADER):
    """Every function the header declares (used by the symbol-export test)."""
    text = re.sub(r"/\*.*?\*/", "", open(header).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(mvd_[a-z0-9_]+)\s*\(", text)))


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "movedepth_b200: %s is missing -- the CUDA extension is required (no CPU/eager "
                "fallback exists). Build it with `python -m movedepth_b200.build`." % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name in declared_symbols():
            if not hasattr(handle, name):
                raise RuntimeError("movedepth_b200: %s does not export %s" % (LIB_PATH, name))
            fn = getattr(handle, name)
            args, res = SIGNATURES[name]
            fn.argtypes = args
            fn.restype = res
        if handle.mvd_version() != 1:
            raise RuntimeError("movedepth_b200: ABI version mismatch")
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().mvd_last_error_string()
        raise RuntimeError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))
