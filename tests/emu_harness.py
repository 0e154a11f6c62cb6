"""TEST INFRASTRUCTURE ONLY -- builds and drives tests/emu/emu.cpp, the CPU
emulation of the fused loss kernels (same phase functions, same tiles and
rings as the CUDA kernels, executed sequentially)."""
import ctypes as C
import os
import subprocess

import torch

from uncertainty_model_b200._lib import UslLossConfig, UslLossScale
from uncertainty_model_b200.functional import (LossSettings, make_config,
                                               make_scale)

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'emu', 'emu.cpp')
OUT = os.path.join(HERE, 'emu', 'build', 'libemu.so')
SRC_COL = os.path.join(HERE, 'emu', 'emu_col.cpp')
OUT_COL = os.path.join(HERE, 'emu', 'build', 'libemu_col.so')
CSRC = os.path.join(os.path.dirname(HERE), 'uncertainty_model_b200', 'csrc')

_emu = None


def emu():
    global _emu
    if _emu is None:
        deps = [SRC] + [os.path.join(CSRC, f) for f in
                        ('loss_core.cuh', 'cons_core.cuh', 'usl_math.cuh')]
        if not os.path.exists(OUT) or any(
                os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
            os.makedirs(os.path.dirname(OUT), exist_ok=True)
            subprocess.run(['g++', '-O2', '-std=c++17', '-shared', '-fPIC',
                            '-ffp-contract=off', '-x', 'c++', SRC, '-o', OUT],
                           check=True)
        lib = C.CDLL(OUT)
        lib.emu_loss_fwd.restype = C.c_int
        lib.emu_loss_fwd.argtypes = [C.POINTER(UslLossConfig),
                                     C.POINTER(UslLossScale), C.c_int, C.c_int,
                                     C.POINTER(C.c_double)]
        lib.emu_cons_scatter.restype = C.c_int
        lib.emu_cons_scatter.argtypes = [C.POINTER(UslLossConfig),
                                         C.POINTER(UslLossScale), C.c_int,
                                         C.POINTER(C.c_float)]
        lib.emu_cons_scatter_add.restype = C.c_int
        lib.emu_cons_scatter_add.argtypes = lib.emu_cons_scatter.argtypes
        lib.emu_loss_bwd.restype = C.c_int
        lib.emu_loss_bwd.argtypes = [C.POINTER(UslLossConfig),
                                     C.POINTER(UslLossScale), C.c_int, C.c_int,
                                     C.c_int, C.POINTER(C.c_float)]
        _emu = lib
    return _emu


def emu_scale(settings: LossSettings, terms, coefs, images, pred, *, recon=None,
              err=None, g=(1.0, 1.0), TW=256, R=32, consR=16,
              want_recon=False, grad_recon_in=None, backward=True):
    """Run the emulated forward (+ backward) of ONE scale on CPU tensors.

    Returns dict(sums[6], err (B,2,h,w), recon, grad_pred, grad_recon)."""
    L = emu()
    b, _, h, w = pred.shape
    images = images.contiguous() if images is not None else None
    pred = pred.contiguous()
    cfg = make_config(terms, settings, coefs)
    err_out = torch.full((b, 2, h, w), float('nan'))
    recon_out = torch.full((b, 6, h, w), float('nan')) if want_recon else None
    sc = make_scale(images, pred[:, 0:2], pred[:, 2:4], shape=(b, h, w),
                    recon_in=recon, err_in=err, err_out=err_out,
                    recon_out=recon_out)
    sums = (C.c_double * 6)()
    L.emu_loss_fwd(C.byref(cfg), C.byref(sc), TW, R, sums)
    out = dict(sums=list(sums), err=err_out, recon=recon_out)
    if backward:
        grad_pred = torch.full((b, 4, h, w), float('nan'))
        grad_recon = torch.full((b, 6, h, w), float('nan')) \
            if recon is not None else None
        sc = make_scale(images, pred[:, 0:2], pred[:, 2:4], shape=(b, h, w),
                        recon_in=recon, err_in=err,
                        grad_recon_in=grad_recon_in,
                        grad_disp=grad_pred[:, 0:2], grad_unc=grad_pred[:, 2:4],
                        grad_recon_out=grad_recon)
        L.emu_loss_bwd(C.byref(cfg), C.byref(sc), TW, R, consR,
                       (C.c_float * 2)(*g))
        out.update(grad_pred=grad_pred, grad_recon=grad_recon)
    return out


_emu_col = None


def emu_col():
    global _emu_col
    if _emu_col is None:
        deps = [SRC_COL] + [os.path.join(CSRC, f) for f in
                            ('col_core.cuh', 'loss_core.cuh', 'usl_math.cuh')]
        if not os.path.exists(OUT_COL) or any(
                os.path.getmtime(d) > os.path.getmtime(OUT_COL) for d in deps):
            os.makedirs(os.path.dirname(OUT_COL), exist_ok=True)
            subprocess.run(['g++', '-O2', '-std=c++17', '-shared', '-fPIC',
                            '-ffp-contract=off', '-x', 'c++', SRC_COL, '-o',
                            OUT_COL], check=True)
        lib = C.CDLL(OUT_COL)
        lib.emu_col_fwd.restype = C.c_int
        lib.emu_col_fwd.argtypes = [C.POINTER(UslLossConfig),
                                    C.POINTER(UslLossScale), C.c_int, C.c_int,
                                    C.POINTER(C.c_double)]
        lib.emu_col_grad.restype = C.c_int
        lib.emu_col_grad.argtypes = [C.POINTER(UslLossConfig),
                                     C.POINTER(UslLossScale), C.c_int, C.c_int,
                                     C.c_int, C.POINTER(C.c_float),
                                     C.POINTER(C.c_double)]
        _emu_col = lib
    return _emu_col


def emu_col_scale(settings: LossSettings, terms, coefs, images, pred, *,
                  g=(1.0, 1.0), maxT=512, R=16, consR=16, want_recon=False,
                  grad_recon_in=None):
    """The column-marching kernels (forward-only mode, then the one-pass
    sums+gradient mode, then the emulated scatter on top) for ONE scale.

    Returns dict(sums, sums_grad, err, recon, grad_pred)."""
    L = emu_col()
    b, _, h, w = pred.shape
    images = images.contiguous()
    pred = pred.contiguous()
    cfg = make_config(terms, settings, coefs)
    err_out = torch.full((b, 2, h, w), float('nan'))
    recon_out = torch.full((b, 6, h, w), float('nan')) if want_recon else None
    sc = make_scale(images, pred[:, 0:2], pred[:, 2:4], shape=(b, h, w),
                    err_out=err_out, recon_out=recon_out)
    sums = (C.c_double * 6)()
    assert L.emu_col_fwd(C.byref(cfg), C.byref(sc), maxT, R, sums) == 0
    grad_pred = torch.full((b, 4, h, w), float('nan'))
    sc = make_scale(images, pred[:, 0:2], pred[:, 2:4], shape=(b, h, w),
                    grad_recon_in=grad_recon_in, grad_disp=grad_pred[:, 0:2],
                    grad_unc=grad_pred[:, 2:4])
    gout = (C.c_float * 2)(*g)
    sums_g = (C.c_double * 6)()
    assert L.emu_col_grad(C.byref(cfg), C.byref(sc), maxT, R, 0, gout,
                          sums_g) == 0
    if terms & 34:          # TERM_CONS_D | TERM_CONS_U: the scatter adds its part
        emu().emu_cons_scatter_add(C.byref(cfg), C.byref(sc), consR, gout)
    return dict(sums=list(sums), sums_grad=list(sums_g), err=err_out,
                recon=recon_out, grad_pred=grad_pred)
