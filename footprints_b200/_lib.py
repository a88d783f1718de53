"""ctypes binding of ``lib/libfootprints_b200.so`` (C ABI declared in ``include/footprints_b200.h``).

There is no CPU / eager fallback: if the shared library is missing or the device is not sm_100 the
product path raises immediately.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libfootprints_b200.so')

FP_F16, FP_BF16 = 0, 1
FP_ACT_NONE, FP_ACT_RELU, FP_ACT_ELU, FP_ACT_SIGMOID = 0, 1, 2, 3
FP_PAD_ZERO, FP_PAD_REFLECT = 0, 1

_i32 = ctypes.c_int32
_i64 = ctypes.c_longlong
_f32 = ctypes.c_float
_vp = ctypes.c_void_p


class FpConvDesc(ctypes.Structure):
    _fields_ = [(n, _i32) for n in (
        'batch', 'c0', 'h0', 'w0', 'up0', 'c1', 'kh', 'kw', 'stride', 'pad_t', 'pad_l', 'pad_mode',
        'ho', 'wo', 'cout', 'kpad', 'npad', 'bn', 'act', 'dtype', 'out_f32', 'out_ld', 'res_ld',
        'b_via_tma', 'w_dtype_plus1')]


class FpSplitLow(ctypes.Structure):
    _fields_ = [('src0_lo', _vp), ('src1_lo', _vp), ('residual_lo', _vp), ('out_lo', _vp)]


class FpPackJob(ctypes.Structure):
    _fields_ = [('w', _vp), ('out', _vp)] + [(n, _i32) for n in (
        'cout', 'cin', 'kh', 'kw', 'cpad', 'npad', 'kpad', 'dtype', 'flip', 'stem')] + [('first', ctypes.c_int64)]


# name -> (restype, argtypes); every name must be exported by the library (tests check this against
# the header as well)
PROTOTYPES = {
    'fp_version': (ctypes.c_int, []),
    'fp_last_error': (ctypes.c_char_p, []),
    'fp_check_device': (ctypes.c_int, []),
    'fp_reload_options': (ctypes.c_int, []),
    'fp_conv_fwd': (ctypes.c_int, [ctypes.POINTER(FpConvDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'fp_conv_fwd_split': (ctypes.c_int, [ctypes.POINTER(FpConvDesc), _vp, _vp, ctypes.POINTER(FpSplitLow), _vp, _vp, _vp, _vp, _vp, _vp]),
    'fp_pack_conv_weight_split': (ctypes.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    'fp_pack_stem_weight_split': (ctypes.c_int, [_vp, _i32, _i32, _i32, _vp, _vp]),
    'fp_stem_prep_split': (ctypes.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    'fp_maxpool3x3s2_split': (ctypes.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    'fp_pack_conv_weight': (ctypes.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    'fp_pack_stem_weight': (ctypes.c_int, [_vp, _i32, _i32, _i32, _vp, _vp]),
    'fp_pack_weights_batched': (ctypes.c_int, [_vp, _i32, _i64, _vp]),
    'fp_stem_prep': (ctypes.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    'fp_maxpool3x3s2': (ctypes.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    'fp_heads_to_nchw': (ctypes.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    'fp_heads_to_nchw_f16_post': (ctypes.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    'fp_psp_fwd': (ctypes.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    'fp_resize_u8_horizontal': (ctypes.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _vp, _vp]),
    'fp_resize_u8_vertical_to_f32': (ctypes.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _vp, _vp]),
    'fp_head_conv_fwd': (ctypes.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    'fp_head_conv_dgrad': (ctypes.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _vp, _vp]),
    'fp_head_conv_wgrad_workspace_floats': (ctypes.c_longlong, [_i32]),
    'fp_head_conv_wgrad': (ctypes.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    'fp_reduce_workspace_floats': (ctypes.c_int, [_i32]),
    'fp_bn_stats': (ctypes.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _f32, _f32, _vp, _vp, _vp, _vp, _vp, _vp]),
    'fp_bn_stats_split': (ctypes.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _f32, _f32, _vp, _vp, _vp, _vp, _vp, _vp]),
    'fp_bn_apply_split': (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _i32, _i32, _vp, _vp, _vp]),
    'fp_bn_apply': (ctypes.c_int, [_vp, _vp, _vp, _vp, _i32, _i64, _i32, _i32, _vp, _vp]),
    'fp_bn_bwd': (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    'fp_bias_grad': (ctypes.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _vp]),
    'fp_loss_fwd_bwd': (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _f32, _f32, _f32, _vp, _vp, _vp]),
    'fp_heads_bwd': (ctypes.c_int, [_vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _f32, _vp, _vp]),
    'fp_fold_bwd': (ctypes.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _vp, _vp]),
    'fp_act_bwd': (ctypes.c_int, [_vp, _vp, _i32, _i64, _i32, _i32, _vp, _vp]),
    'fp_grad_add': (ctypes.c_int, [_vp, _vp, _i64, _i32, _vp]),
    'fp_maxpool3x3s2_bwd': (ctypes.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    'fp_conv_wgrad_workspace_bytes': (ctypes.c_longlong, [ctypes.POINTER(FpConvDesc)]),
    'fp_conv_wgrad': (ctypes.c_int, [ctypes.POINTER(FpConvDesc), _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    'fp_debug_read_trace': (ctypes.c_int, [_vp]),
    'fp_adam_step': (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _i32, _f32, _vp]),
    'fp_adam_step_dev': (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _vp]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f'footprints_b200: {LIB_PATH} is missing. Build it with `python -c "import '
            f'__graft_entry__ as g; g.build()"` or `make -C footprints_b200/csrc`. '
            f'There is no CPU fallback.')
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)       # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise RuntimeError(f'footprints_b200 error {rc}: {load().fp_last_error().decode()}')


def require_device():
    """Raise unless the current CUDA device is a B200 (sm_100)."""
    check(load().fp_check_device())
