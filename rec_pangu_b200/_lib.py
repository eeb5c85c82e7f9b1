"""ctypes binding of librec_pangu_b200.so (the C ABI declared in include/rec_pangu_b200.h).

There is no fallback: if the library cannot be loaded (and cannot be built with nvcc), importing the ops raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'librec_pangu_b200.so')
ABI_VERSION = 8
MAX_FIELDS = 64
MAX_DENSE = 64
ERR_UNSUPPORTED = -1
ADAM_MAX_TENSORS = 32

_vp = C.c_void_p
_i32 = C.c_int32
_i64 = C.c_int64
_f32 = C.c_float


class GatherDesc(C.Structure):
    _fields_ = [('B', _i32), ('F', _i32), ('D', _i32), ('Nd', _i32), ('ldx', _i32), ('ld_lr', _i32),
                ('tables', C.POINTER(_vp)), ('rows', C.POINTER(_i64)), ('idx', C.POINTER(_vp)),
                ('dense', C.POINTER(_vp)), ('lr_tables', C.POINTER(_vp)),
                ('x', _vp), ('fm', _vp), ('fm_s', _vp), ('lr_in', _vp), ('err', _vp), ('G', _i32), ('shard_tab', _vp)]


class ScatterDesc(C.Structure):
    _fields_ = [('B', _i32), ('F', _i32), ('D', _i32), ('lddx', _i32), ('ldx', _i32), ('ld_dlr', _i32),
                ('grads', C.POINTER(_vp)), ('lr_grads', C.POINTER(_vp)), ('rows', C.POINTER(_i64)),
                ('idx', C.POINTER(_vp)), ('dx', _vp), ('x', _vp), ('dfm', _vp), ('fm_s', _vp), ('dlr_in', _vp),
                ('G', _i32), ('grad_shard_tab', _vp)]


class SparseAdamDesc(C.Structure):
    _fields_ = [('B', _i32), ('F', _i32), ('D', _i32), ('step', _i32), ('lr', _f32), ('beta1', _f32), ('beta2', _f32),
                ('eps', _f32), ('weights', C.POINTER(_vp)), ('grads', C.POINTER(_vp)), ('exp_avg', C.POINTER(_vp)),
                ('exp_avg_sq', C.POINTER(_vp)), ('stamps', C.POINTER(_vp)), ('rows', C.POINTER(_i64)),
                ('idx', C.POINTER(_vp)), ('step_dev', _vp)]


class AdamMultiDesc(C.Structure):
    _fields_ = [('count', _i32), ('step', _i32), ('lr', _f32), ('beta1', _f32), ('beta2', _f32), ('eps', _f32),
                ('params', C.POINTER(_vp)), ('grads', C.POINTER(_vp)), ('exp_avg', C.POINTER(_vp)),
                ('exp_avg_sq', C.POINTER(_vp)), ('numel', C.POINTER(_i64)), ('step_dev', _vp)]


class TowerFwdDesc(C.Structure):
    _fields_ = [('M', _i32), ('H', _i32), ('n_tail', _i32), ('h1', _vp), ('ldh1', _i64),
                ('W', C.POINTER(_vp)), ('b', C.POINTER(_vp)), ('h', C.POINTER(_vp)),
                ('w_out', _vp), ('b_out', _vp), ('addend', _vp), ('logit', _vp), ('label', _vp), ('pred', _vp),
                ('loss', _vp), ('eps', _f32), ('scale', _f32), ('work', _vp)]


class TowerBwdDesc(C.Structure):
    _fields_ = [('M', _i32), ('H', _i32), ('n_tail', _i32), ('hin', C.POINTER(_vp)), ('ldh1', _i64),
                ('W', C.POINTER(_vp)), ('w_out', _vp), ('dz', C.POINTER(_vp)), ('db', C.POINTER(_vp)),
                ('dw_out', _vp), ('db_out', _vp), ('pred', _vp), ('label', _vp), ('gloss', _vp),
                ('eps', _f32), ('scale', _f32), ('dlogit_in', _vp), ('dlogit_out', _vp)]


# name -> (restype, argtypes); must list every symbol of include/rec_pangu_b200.h (tests check this)
SIGNATURES = {
    'rpb_version': (C.c_int, []),
    'rpb_error_string': (C.c_char_p, [C.c_int]),
    'rpb_set_option': (C.c_int, [C.c_char_p, _i64]),
    'rpb_debug_tc_trace': (C.c_int, [C.POINTER(C.c_uint64), C.c_int]),
    'rpb_gather_fwd': (C.c_int, [C.POINTER(GatherDesc), _vp]),
    'rpb_gather_bwd': (C.c_int, [C.POINTER(ScatterDesc), _vp]),
    'rpb_rows_zero': (C.c_int, [C.POINTER(ScatterDesc), _vp]),
    'rpb_hash_to_row': (C.c_int, [_vp, _vp, _i64, _i64, _vp]),
    'rpb_fm_fwd': (C.c_int, [_vp, _i64, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]),
    'rpb_fm_bwd': (C.c_int, [_vp, _i64, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _i64, C.c_int, _vp]),
    'rpb_linear_fwd': (C.c_int, [_vp, _i64, _vp, _vp, _vp, _i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    'rpb_linear_bwd': (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _vp, _i64, _vp, _vp,
                                 C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    'rpb_linear_dx_scatter': (C.c_int, [_vp, _i64, _vp, C.c_int, C.c_int, C.c_int, C.POINTER(ScatterDesc), _vp]),
    'rpb_rowdot_fwd': (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp]),
    'rpb_rowdot_bwd': (C.c_int, [_vp, _vp, _i64, _vp, _vp, _i64, _vp, _i64, _vp, _vp, C.c_int, C.c_int, _vp]),
    'rpb_sigmoid_bce_fwd': (C.c_int, [_vp, _vp, _vp, _vp, _f32, _f32, C.c_int, _vp, _vp]),
    'rpb_sigmoid_bce_bwd': (C.c_int, [_vp, _vp, _vp, _f32, _f32, _vp, C.c_int, _vp]),
    'rpb_essm_head_fwd': (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _f32, C.c_int, _vp, _vp]),
    'rpb_essm_head_bwd': (C.c_int, [_vp, _vp, _vp, _vp, _vp, _f32, _vp, _vp, C.c_int, _vp]),
    'rpb_tower_tail_fwd': (C.c_int, [C.POINTER(TowerFwdDesc), _vp]),
    'rpb_linear_tower_fwd': (C.c_int, [_vp, _i64, _vp, _vp, C.c_int, C.POINTER(TowerFwdDesc), _vp]),
    'rpb_deepfm_fwd_fused': (C.c_int, [C.POINTER(GatherDesc), _vp, _vp, C.POINTER(TowerFwdDesc), _vp]),
    'rpb_debug_fused_trace': (C.c_int, [C.POINTER(C.c_uint64), C.c_int]),
    'rpb_debug_fused_cta_times': (C.c_int, [_vp]),
    'rpb_tower_tail_bwd': (C.c_int, [C.POINTER(TowerBwdDesc), _vp]),
    'rpb_layernorm_fwd': (C.c_int, [_vp, _i64, _vp, _vp, _f32, _vp, _i64, _vp, _vp, _i32, _i32, _vp]),
    'rpb_layernorm_bwd': (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i32, _i32, _vp]),
    'rpb_dropout_fwd': (C.c_int, [_vp, _vp, _i64, _f32, C.c_uint64, _vp, _vp]),
    'rpb_dropout_bwd': (C.c_int, [_vp, _vp, _vp, _i64, _f32, C.c_uint64, _vp, _vp]),
    'rpb_crossnet_fwd': (C.c_int, [_vp, _i64, C.c_int, C.c_int, C.POINTER(_vp), C.POINTER(_vp), _vp, _i64, _vp, C.c_int, _vp]),
    'rpb_crossnet_bwd': (C.c_int, [_vp, _i64, C.c_int, C.c_int, C.POINTER(_vp), C.POINTER(_vp), _vp, _vp, _i64, _vp, _i64,
                                   C.POINTER(_vp), C.POINTER(_vp), C.c_int, _vp]),
    'rpb_cin_fwd': (C.c_int, [_vp, _i64, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_i32), C.POINTER(_vp),
                              C.POINTER(_vp), _vp, _i64, _vp]),
    'rpb_cin_bwd': (C.c_int, [_vp, _i64, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_i32), C.POINTER(_vp),
                              C.POINTER(_vp), _vp, _i64, _vp, _i64, C.c_int, C.POINTER(_vp), C.POINTER(_vp), _vp]),
    'rpb_cin_fwd_save': (C.c_int, [_vp, _i64, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_i32), C.POINTER(_vp),
                                   C.POINTER(_vp), _vp, _i64, _vp, _i64, _vp]),
    'rpb_cin_bwd_saved': (C.c_int, [_vp, _i64, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_i32), C.POINTER(_vp),
                                    C.POINTER(_vp), _vp, _i64, _vp, _i64, C.c_int, C.POINTER(_vp), C.POINTER(_vp), _vp, _i64, _vp]),
    'rpb_matmul_kn_fwd': (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp, _i64, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    'rpb_matmul_kn_bwd': (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, C.c_int, C.c_int, C.c_int,
                                    C.c_int, _vp]),
    'rpb_mmoe_combine_fwd': (C.c_int, [_vp, _i64, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]),
    'rpb_mmoe_combine_bwd': (C.c_int, [_vp, _i64, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _i64, _vp]),
    'rpb_bn_stats': (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    'rpb_bn_apply': (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp]),
    'rpb_bn_bwd': (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp]),
    'rpb_bn_bwd_stats': (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp]),
    'rpb_bn_bwd_dx': (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, _f32, C.c_int, _vp]),
    'rpb_fibinet_fwd': (C.c_int, [_vp, _i64, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_int, _vp, _vp, _vp, _i64, _vp, _vp]),
    'rpb_fibinet_bwd': (C.c_int, [_vp, _i64, C.c_int, C.c_int, C.c_int, _vp, C.c_int, _vp, _vp, _vp, _vp, _i64, _vp, _i64,
                                  _vp, _vp, _vp, _vp]),
    'rpb_adam_dense': (C.c_int, [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, C.c_int, _vp]),
    'rpb_adam_multi': (C.c_int, [C.POINTER(AdamMultiDesc), _vp]),
    'rpb_sparse_adam': (C.c_int, [C.POINTER(SparseAdamDesc), _vp]),
    'rpb_sparse_adam_catchup': (C.c_int, [C.POINTER(SparseAdamDesc), _vp]),
    'rpb_sparse_adam_flush': (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _f32, _f32, _f32, _f32, _vp, _i32, _vp]),
    'rpb_autoint_attn_fwd': (C.c_int, [_vp, _i64, _vp, _i64, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    'rpb_autoint_attn_bwd': (C.c_int, [_vp, _i64, C.c_int, _vp, _vp, _vp, _i64, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
}

_lib = None


class RpbError(RuntimeError):
    pass


def load():
    """Load (building first if the .so is absent and nvcc exists).  Raises loudly otherwise."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import build as _build
        _build.build(verbose=False)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    v = lib.rpb_version()
    if v != ABI_VERSION:
        raise RpbError(f'librec_pangu_b200.so ABI version {v} != expected {ABI_VERSION}; rebuild with '
                       f'`python -m rec_pangu_b200.build --force`')
    _lib = lib
    return lib


def check(code, what):
    if code != 0:
        msg = load().rpb_error_string(code).decode()
        raise RpbError(f'{what} failed: {msg} (code {code})')
