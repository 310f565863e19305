"""ctypes binding of libekl_b200.so (the C ABI declared in include/ekl_b200.h).

The product path has no fallback: if the shared library is missing or the device is not sm_100a the ops layer
fails loudly.  `python -m text2img_ekl_b200.build` (or __graft_entry__.build()) compiles it in-tree.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EKL_LIB_PATH") or os.path.join(HERE, "libekl_b200.so")

S1, UP2, DOWN2 = 0, 1, 2
IMPL_TC, IMPL_SIMT = 0, 1
FMT_NHWC_BF16, FMT_NCHW_F32 = 0, 1
ACT_NONE, ACT_GLU, ACT_LRELU, ACT_RELU, ACT_TANH = 0, 1, 2, 3, 4
W_KRSC, W_KCRS = 0, 1


class EklConv(C.Structure):
    _fields_ = [("mode", C.c_int), ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Cin", C.c_int), ("Cout", C.c_int),
                ("group_b", C.c_int), ("impl", C.c_int), ("x_fmt", C.c_int), ("y_fmt", C.c_int), ("act", C.c_int), ("w_layout", C.c_int),
                ("w_cin_total", C.c_int), ("w_cin_off", C.c_int), ("w_cout_valid", C.c_int)]


class EklError(RuntimeError):
    pass


_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float
_cp = C.POINTER(EklConv)
_ip = C.POINTER(C.c_int)

# name -> (restype, argtypes); every symbol include/ekl_b200.h declares
SIGNATURES = {
    "ekl_last_error": (C.c_char_p, []),
    "ekl_version": (_i, []),
    "ekl_require_sm100": (_i, []),
    "ekl_conv_packed_elems": (_i64, [_cp, _i]),
    "ekl_conv_pack": (_i, [_cp, _vp, _vp, _vp, _vp]),
    "ekl_conv_fwd": (_i, [_cp, _vp, _vp, _vp, _vp, _vp]),
    "ekl_conv_workspace_elems": (_i64, [_cp, _i]),
    "ekl_conv_fwd_ws": (_i, [_cp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ekl_conv_bwd_data_ws": (_i, [_cp, _vp, _vp, _vp, _vp, _vp]),
    "ekl_conv_dgrad_from_fwd": (_i, [_cp]),
    "ekl_conv_route": (_i, [_cp, _i]),
    "ekl_conv_bwd_data_fw": (_i, [_cp, _vp, _vp, _vp, _vp, _vp]),
    "ekl_conv_fwd_bias9": (_i, [_cp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ekl_conv_bwd_data": (_i, [_cp, _vp, _vp, _vp, _vp]),
    "ekl_conv_split_bn_fusable": (_i, [_cp, _i]),
    "ekl_conv_split_bn_aux_floats": (_i64, [_cp]),
    "ekl_conv_fwd_split_bn_act": (_i, [_cp, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "ekl_conv_bwd_weight": (_i, [_cp, _vp, _vp, _vp, _vp]),
    "ekl_conv_plan_dump": (_i, [_cp, _i, C.POINTER(C.c_int), _i]),
    "ekl_col_stats": (_i, [_vp, _i64, _i, _i, _vp, _vp]),
    "ekl_code_bias9_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "ekl_border_sums9": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "ekl_code_bias9_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "ekl_bn_act_fwd": (_i, [_vp, _i64, _i, _i, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "ekl_bn_bwd_scratch_doubles": (_i64, [_i64, _i, _i, _i]),
    "ekl_bn_act_bwd": (_i, [_vp, _vp, _i64, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "ekl_lrelu_bwd": (_i, [_vp, _vp, _vp, _i64, _vp]),
    "ekl_cat_code": (_i, [_vp, _i, _vp, _i, _i, _i, _vp, _vp]),
    "ekl_cat_code_bwd": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "ekl_img_s2d": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "ekl_img_s2d_bwd": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "ekl_img_pyramid_level": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    "ekl_head_tanh_fwd": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "ekl_head_tanh_bwd": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "ekl_color_stats_fwd": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp]),
    "ekl_color_stats_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "ekl_dhead_dots": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "ekl_dhead_dots_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ekl_dloss_fwd": (_i, [_i, _i, _i, _ip, _ip, _ip, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ekl_reparam_kl_fwd": (_i, [_vp, _i64, _vp, _i64, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "ekl_reparam_kl_bwd": (_i, [_vp, _i64, _vp, _i64, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ekl_linear_bn_relu_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _f, _i, _vp, _vp, _vp, _vp]),
    "ekl_linear_bn_relu_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ekl_caps_supported": (_i, [_i, _i, _i, _i]),
    "ekl_caps_proj_u": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "ekl_caps_proj_s": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "ekl_caps_squash_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "ekl_caps_outer": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "ekl_caps_agree_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "ekl_caps_agree_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "ekl_caps_route_supported": (_i, [_i, _i, _i, _i]),
    "ekl_caps_route_fwd": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "ekl_caps_route_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "ekl_adam_step": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _f, _f, _f, _f, _vp]),
    "ekl_adam_step_g16": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _f, _f, _f, _f, _vp]),
    "ekl_adam_tick": (_i, [_vp, _f, _f, _vp]),
    "ekl_adam_apply": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i64, _vp, _f, _f, _f, _f, _vp]),
    "ekl_cast_bf16": (_i, [_vp, _vp, _i64, _vp]),
    "ekl_dloss_bwd": (_i, [_i, _i, _i, _ip, _ip, _ip, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EklError("libekl_b200.so not built (%s). Run `python -m text2img_ekl_b200.build`; there is no "
                           "fallback path." % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise EklError("ekl_b200 error %d: %s" % (rc, lib().ekl_last_error().decode()))


def ptr(t):
    """device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
