"""ctypes binding of libusl.so (the C ABI declared in include/usl.h).

There is no CPU fallback: if the shared library is missing the import of any
op fails loudly, and every entry point raises on a non-zero return code.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libusl.so')

USL_NUM_TERMS = 6
USL_MAX_SCALES = 8
USL_ERR_UNSUPPORTED = -3
USL_SCALE_GENERAL_KERNELS = 1
USL_MAX_RESCALE = 16

TERM_REPROJ, TERM_CONS_D, TERM_SMOOTH_D = 1, 2, 4
TERM_UNC, TERM_SMOOTH_U, TERM_CONS_U = 8, 16, 32
LOSS_TYPES = {'l1': 0, 'bayesian': 1, 'log_bayesian': 2}

_f32p = C.c_void_p      # device pointers travel as integers


class UslLossConfig(C.Structure):
    _fields_ = [('terms', C.c_uint32), ('loss_type', C.c_int32),
                ('alpha', C.c_float), ('c1', C.c_float), ('c2', C.c_float),
                ('coef', C.c_float * USL_NUM_TERMS)]


class UslLossScale(C.Structure):
    _fields_ = [
        ('B', C.c_int32), ('h', C.c_int32), ('w', C.c_int32),
        ('flags', C.c_int32),
        ('images', _f32p), ('img_bs', C.c_int64), ('img_cs', C.c_int64),
        ('disp', _f32p), ('disp_bs', C.c_int64), ('disp_cs', C.c_int64),
        ('unc', _f32p), ('unc_bs', C.c_int64), ('unc_cs', C.c_int64),
        ('recon_in', _f32p), ('rin_bs', C.c_int64), ('rin_cs', C.c_int64),
        ('err_in', _f32p), ('ein_bs', C.c_int64), ('ein_cs', C.c_int64),
        ('recon_out', _f32p), ('err_out', _f32p),
        ('grad_recon_in', _f32p),
        ('grad_disp', _f32p), ('gd_bs', C.c_int64), ('gd_cs', C.c_int64),
        ('grad_unc', _f32p), ('gu_bs', C.c_int64), ('gu_cs', C.c_int64),
        ('grad_recon_out', _f32p),
        ('scatter_ws', _f32p),
    ]


class UslDiscLevel(C.Structure):
    _fields_ = [
        ('B', C.c_int32), ('h', C.c_int32), ('w', C.c_int32),
        ('reserved', C.c_int32),
        ('images', _f32p), ('img_bs', C.c_int64), ('img_cs', C.c_int64),
        ('pred', _f32p), ('pred_bs', C.c_int64), ('pred_cs', C.c_int64),
        ('recon', _f32p), ('rec_bs', C.c_int64), ('rec_cs', C.c_int64),
        ('out', _f32p),
    ]


# name -> (restype, argtypes); mirrors include/usl.h one to one
SIGNATURES = {
    'usl_version': (C.c_int, []),
    'usl_strerror': (C.c_char_p, [C.c_int]),
    'usl_launch_count': (C.c_longlong, []),
    'usl_pyramid': (C.c_int, [_f32p, C.c_int, C.c_int, C.c_int, C.c_int,
                              C.c_longlong, C.c_longlong, C.c_int,
                              C.POINTER(C.c_void_p), C.c_void_p]),
    'usl_warp_fwd': (C.c_int, [_f32p, C.c_longlong, C.c_float, _f32p,
                               C.c_longlong, C.c_longlong, C.c_int, C.c_int,
                               C.c_int, C.c_int, _f32p, C.c_longlong,
                               C.c_longlong, C.c_void_p]),
    'usl_warp_bwd_disp': (C.c_int, [_f32p, C.c_longlong, C.c_float, _f32p,
                                    C.c_longlong, C.c_longlong, _f32p,
                                    C.c_longlong, C.c_longlong, C.c_int,
                                    C.c_int, C.c_int, C.c_int, _f32p,
                                    C.c_longlong, C.c_void_p]),
    'usl_warp_bwd_image_workspace_bytes': (C.c_longlong, [C.c_int, C.c_int,
                                                          C.c_int, C.c_int]),
    'usl_warp_bwd_image': (C.c_int, [_f32p, C.c_longlong, C.c_float, _f32p,
                                     C.c_longlong, C.c_longlong, C.c_int,
                                     C.c_int, C.c_int, C.c_int, C.c_void_p,
                                     _f32p, C.c_longlong, C.c_longlong,
                                     C.c_void_p]),
    'usl_loss_plan': (C.c_int, [C.POINTER(UslLossConfig),
                                C.POINTER(UslLossScale), C.c_int, C.c_int,
                                C.POINTER(C.c_int)]),
    'usl_loss_grad': (C.c_int, [C.POINTER(UslLossConfig),
                                C.POINTER(UslLossScale), C.c_int, _f32p, _f32p,
                                _f32p, C.c_int, C.c_void_p]),
    'usl_loss_grad_sharded': (C.c_int, [C.POINTER(UslLossConfig),
                                        C.POINTER(UslLossScale), C.c_int,
                                        _f32p, C.POINTER(C.c_int), C.c_void_p,
                                        C.c_void_p, C.c_void_p]),
    'usl_debug_timeline': (C.c_int, [C.c_void_p]),
    'usl_loss_fwd_ctas': (C.c_int, [C.POINTER(UslLossScale)]),
    'usl_loss_fwd': (C.c_int, [C.POINTER(UslLossConfig),
                               C.POINTER(UslLossScale), C.c_int, _f32p,
                               C.c_void_p]),
    'usl_loss_reduce': (C.c_int, [_f32p, C.POINTER(C.c_int), C.c_int,
                                  C.c_void_p, C.c_void_p]),
    'usl_loss_combine': (C.c_int, [C.c_void_p, _f32p, C.c_int, _f32p, _f32p,
                                   C.c_void_p]),
    'usl_loss_bwd': (C.c_int, [C.POINTER(UslLossConfig),
                               C.POINTER(UslLossScale), C.c_int, _f32p, _f32p,
                               C.c_int, C.c_void_p]),
    'usl_grad_rescale': (C.c_int, [_f32p, C.POINTER(C.c_void_p),
                                   C.POINTER(C.c_longlong), C.c_int,
                                   C.c_void_p]),
    'usl_head_fwd': (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                               C.POINTER(C.c_longlong), C.c_int, C.c_float,
                               C.c_void_p]),
    'usl_head_bwd': (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                               C.POINTER(C.c_void_p), C.POINTER(C.c_longlong),
                               C.c_int, C.c_float, C.c_void_p]),
    'usl_disc_input': (C.c_int, [C.POINTER(UslDiscLevel), C.c_int,
                                 C.c_void_p]),
    'usl_recon_bwd': (C.c_int, [C.POINTER(UslDiscLevel), C.POINTER(C.c_void_p),
                                C.POINTER(C.c_void_p), C.POINTER(C.c_longlong),
                                C.POINTER(C.c_longlong), C.c_int, C.c_int,
                                C.c_void_p]),
    'usl_ssim_workspace_bytes': (C.c_size_t, [C.c_int, C.c_int, C.c_int,
                                              C.c_int, C.c_int]),
    'usl_ssim_gauss': (C.c_int, [_f32p, C.c_longlong, C.c_longlong, _f32p,
                                 C.c_longlong, C.c_longlong, C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.c_int, C.c_float,
                                 C.c_float, C.c_float, C.c_float, _f32p,
                                 C.c_void_p, C.c_size_t, C.c_void_p]),
    'usl_combine_disparity': (C.c_int, [_f32p, _f32p, C.c_int, C.c_int,
                                        C.c_int, C.c_double, C.c_double,
                                        C.c_void_p, C.c_void_p]),
    'usl_heatmap': (C.c_int, [_f32p, C.c_longlong, C.c_int, C.c_void_p,
                              C.c_int, C.c_void_p, C.c_void_p]),
    'usl_pool3_fwd': (C.c_int, [_f32p, C.c_longlong, C.c_longlong, C.c_int,
                                C.c_int, C.c_int, C.c_int, _f32p,
                                C.c_void_p]),
    'usl_pool3_bwd': (C.c_int, [_f32p, C.c_int, C.c_int, C.c_int, C.c_int,
                                _f32p, C.c_void_p]),
    'usl_spars_workspace_bytes': (C.c_size_t, [C.c_int, C.c_int, C.c_int,
                                               C.c_int, C.c_int]),
    'usl_spars_curve': (C.c_int, [_f32p, _f32p, C.c_int, C.c_int, C.c_int,
                                  C.c_int, C.POINTER(C.c_int), C.c_int,
                                  C.c_void_p, C.c_void_p, _f32p, _f32p,
                                  C.c_void_p, C.c_size_t, C.c_void_p]),
    'usl_spars_finish': (C.c_int, [C.c_void_p, C.c_int, C.c_longlong, _f32p,
                                   C.c_void_p]),
    'usl_spars_ause': (C.c_int, [_f32p, _f32p, C.c_int, _f32p, C.c_void_p]),
}

_lib = None


class UslError(RuntimeError):
    pass


def lib():
    """The loaded library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise UslError(
                f'{LIB_PATH} is missing: build it with '
                '`python -m uncertainty_model_b200._build` (needs nvcc). '
                'There is no CPU fallback.')
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError if not exported
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str = '') -> None:
    if rc != 0:
        msg = lib().usl_strerror(rc).decode()
        raise UslError(f'libusl {what}: {msg} (rc={rc})')
