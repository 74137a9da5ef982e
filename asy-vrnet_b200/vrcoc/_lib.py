"""ctypes binding of libvrcoc.so (the C-ABI declared in include/vrcoc.h).

The handle lives at module scope, never on an nn.Module: ModelEMA deep-copies the model
(reference nets/yolo_training.py:457) and nn.DataParallel replicates it (reference yolo.py:103).
There is no fallback: if the library is missing, import of the product package fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VRCOC_LIB", os.path.join(os.path.dirname(_HERE), "csrc", "libvrcoc.so"))

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_GELU, ACT_SILU, ACT_LRELU = 0, 1, 2, 3, 4
ENGINE_AUTO, ENGINE_SIMT, ENGINE_TC = 0, 1, 2


class VrcocError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    """Mirror of `struct vrcoc_conv_desc` (include/vrcoc.h)."""
    _fields_ = [
        ("B", C.c_int32), ("H_in", C.c_int32), ("W_in", C.c_int32), ("H_out", C.c_int32), ("W_out", C.c_int32),
        ("C0", C.c_int32), ("C1", C.c_int32), ("O", C.c_int32),
        ("kh", C.c_int32), ("kw", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32),
        ("src0", C.c_void_p), ("src0_dtype", C.c_int32), ("src0_bstride", C.c_int64),
        ("src1", C.c_void_p), ("src1_dtype", C.c_int32), ("src1_bstride", C.c_int64),
        ("chan_src", C.c_void_p),
        ("gn_sums", C.c_void_p), ("gn_gamma", C.c_void_p), ("gn_beta", C.c_void_p), ("gn_eps", C.c_float),
        ("table", C.c_void_p), ("has_gate", C.c_int32),
        ("weight", C.c_void_p), ("weight_dtype", C.c_int32),
        ("e_scale", C.c_void_p), ("e_shift", C.c_void_p), ("act", C.c_int32), ("post_scale", C.c_void_p),
        ("res", C.c_void_p), ("res_dtype", C.c_int32),
        ("f_scale", C.c_void_p), ("f_shift", C.c_void_p),
        ("out", C.c_void_p), ("out_dtype", C.c_int32),
        ("out2", C.c_void_p), ("out2_dtype", C.c_int32), ("O_split", C.c_int32),
        ("out_sample_sums", C.c_void_p), ("out_minmax", C.c_void_p),
        ("engine", C.c_int32), ("dil", C.c_int32), ("k_order", C.c_int32),
        ("gn_fold_k1", C.c_void_p),
    ]


_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> (restype, argtypes); every symbol include/vrcoc.h declares
SIGNATURES = {
    "vrcoc_version": (C.c_char_p, []),
    "vrcoc_last_error": (C.c_char_p, []),
    "vrcoc_device_ok": (_I, []),
    "vrcoc_channel_sums": (_I, [_P, _I, _I, _I, _I, _P, _P, _P]),
    "vrcoc_conv_fwd": (_I, [C.POINTER(ConvDesc), _P]),
    "vrcoc_gn_fold_supported": (_I, [_I, _I, _I, _I, _I]),
    "vrcoc_table_apply": (_I, [C.POINTER(ConvDesc), _P]),
    "vrcoc_conv1x1_wgrad": (_I, [C.POINTER(ConvDesc), _P, _I, _P, _P, _P, _L, _P]),
    "vrcoc_conv1x1_wgrad_workspace": (_L, [C.POINTER(ConvDesc)]),
    "vrcoc_cluster_core_fwd": (_I, [_P, _I, _P, _I, _P, _I, _P, _P, _P, _P] + [_I] * 9 + [_L] * 3 + [_P]),
    "vrcoc_cluster_core_bwd": (_I, [_P, _I, _P, _I, _P, _I, _P, _P, _P, _P, _P, _I, _P, _I, _P, _P] + [_I] * 9 + [_L] * 5 + [_P]),
    "vrcoc_sa_gate_sums": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "vrcoc_fusion_stats": (_I, [_P, _P, _I, _I, _I, _I, _I, _I] + [_P] * 8 + [_P]),
    "vrcoc_radar_enh_table": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "vrcoc_radar_enh_table_concat_order": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "vrcoc_chan_affine": (_I, [_P, _I, _P, _I, _P, _I, _P, _P, _I, _P, _P, _I, _I, _I, _P, _P, _P]),
    "vrcoc_img_enh_finish": (_I, [_P, _I, _P, _I, _P, _I, _P, _P, _P, _I, _I, _I, _P, _P]),
    "vrcoc_debug_set_trace": (_I, [_P]),
    "vrcoc_debug_set_cm": (_I, [_I]),
    "vrcoc_mlp_fused_supported": (_I, [_I, _I, _I, _I]),
    "vrcoc_mlp_fused_fwd": (_I, [_P, _P, _P, _P, _F, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "vrcoc_debug_set_tm_trace": (_I, [_P]),
    "vrcoc_token_mixer_supported": (_I, [_I] * 10),
    "vrcoc_token_mixer_fwd": (_I, [_P, _P, _F, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P] + [_I] * 8 + [_P]),
    "vrcoc_token_mixer_core_supported": (_I, [_I] * 10),
    "vrcoc_token_mixer_core_fwd": (_I, [_P, _P, _F, _P, _P, _P, _P, _P, _P, _P, _P] + [_I] * 8 + [_P]),
    "vrcoc_im2col": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "vrcoc_im2col_rows": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "vrcoc_patch_embed_supported": (_I, [_I] * 7),
    "vrcoc_patch_embed": (_I, [_P, _P, _L, _P, _P, _P, _P] + [_I] * 8 + [_P]),
    "vrcoc_dwconv": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "vrcoc_upsample_bilinear": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "vrcoc_upsample_argmax_supported": (_I, [_I] * 5),
    "vrcoc_upsample_argmax": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "vrcoc_gn_bwd_coef": (_I, [_P, _P, _P, _F, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "vrcoc_proj_res_bwd_coef": (_I, [_P, _P, _P, _I, _P, _P, _I, _I, _P, _P, _P, _P, _P]),
    "vrcoc_col2im": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "vrcoc_chan_bwd_sums": (_I, [_P, _P, _P, _I, _I, _P, _P, _I, _I, _I, _P, _P]),
    "vrcoc_chan_bwd_apply": (_I, [_P, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "vrcoc_img_enh_bwd": (_I, [_P, _P, _P, _I, _P, _I, _I, _I, _P, _P, _P, _P]),
    "vrcoc_minmax_scatter": (_I, [_P, _P, _I, _P, _P, _L, _P]),
    "vrcoc_dwconv3_wgrad": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _P]),
    "vrcoc_bn_stats": (_I, [_P, _I, _I, C.c_double, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P, _P, _P]),
    "vrcoc_bn_bwd_coef": (_I, [_P, _I, _I, _P, _P, _P, _P, _F, C.c_double, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "vrcoc_decode_outputs": (_I, [_P, _P, _P, _I] + [_I] * 10 + [_P, _P]),
    "vrcoc_upsample_bilinear_bwd": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "vrcoc_table_bwd_sums": (_I, [_P, _P, _I, _P, _P, _I, _I, _I, _I, _P, _P]),
    "vrcoc_table_bwd_apply": (_I, [_P, _P, _P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _P]),
    "vrcoc_gelu_bwd": (_I, [_P, _P, _P, _I, _L, _P]),
    "vrcoc_gn_bwd_sums": (_I, [_P, _P, _I, _I, _I, _I, _P, _P]),
    "vrcoc_gn_bwd_apply": (_I, [_P, _P, _P, _P, _I, _P, _P, _P, _I, _I, _I, _P]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise VrcocError(
            f"libvrcoc.so not found at {LIB_PATH}: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU / PyTorch fallback for the CoC + fusion path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc, what=""):
    if rc != 0:
        msg = lib.vrcoc_last_error().decode("utf-8", "replace")
        raise VrcocError(f"{what or 'vrcoc'} failed (code {rc}): {msg}")


def version():
    return lib.vrcoc_version().decode()
