"""Sparsification curves and AUSE / AURG -- drop-in for the reference's
`train/sparsification.py`.

`curve` runs on the GPU: bit-exact k x k average pooling, a stable segmented
radix sort (order = predicted error descending, ties by ascending index, i.e.
`argsort(descending=True, stable=True)`), canonical fp64 tail sums.  The
reference leaves the order of ties and the summation order unspecified; the
definitions used here are spelled out in csrc/spars.cu.
"""
import ctypes as C
from typing import List, Optional, Tuple

import torch
from torch import Tensor

from .. import functional as K
from .._lib import check, lib
from .utils import Device

# frames are processed in chunks so the sort workspace stays below this
WORKSPACE_BUDGET_BYTES = 6 << 30


def cut_points(n: int, steps: int) -> List[int]:
    """Number of removed pixels per step, exactly as the reference computes
    it in Python float arithmetic (sparsification.py:26-27), plus n."""
    return [int(step / steps * n) for step in range(steps)] + [n]


def curve_sums(oracle_error: Tensor, predicted_error: Tensor,
               kernel_size: int = 11, steps: int = 100,
               return_order: bool = False
               ) -> Tuple[Tensor, int, Optional[dict]]:
    """Sum over the rows (frame, view) of this batch of the normalised tail
    means: fp64[steps] on the device, plus the row count.  This is the
    quantity that is all-reduced when frames are sharded over ranks."""
    K.require_cuda_f32(oracle_error, 'oracle_error')
    K.require_cuda_f32(predicted_error, 'predicted_error')
    if oracle_error.shape != predicted_error.shape or oracle_error.dim() != 4:
        raise ValueError('oracle_error and predicted_error must both be '
                         '(B,2,H,W)')
    o = oracle_error.contiguous()
    p = predicted_error.contiguous()
    b, v, h, w = o.shape
    rows = b * v
    oh, ow = h - kernel_size + 1, w - kernel_size + 1
    if oh < 1 or ow < 1:
        raise ValueError('maps are smaller than the pooling kernel')
    n = oh * ow
    cuts = cut_points(n, steps)
    cuts_c = (C.c_int * (steps + 1))(*cuts)
    L = lib()
    acc = torch.zeros(steps, dtype=torch.float64, device=o.device)
    per_row = max(1, L.usl_spars_workspace_bytes(1, h, w, kernel_size,
                                                 int(return_order)))
    chunk = max(1, min(rows, WORKSPACE_BUDGET_BYTES // per_row))
    need = L.usl_spars_workspace_bytes(chunk, h, w, kernel_size,
                                       int(return_order))
    if need == 0:
        raise ValueError('unsupported sparsification shape')
    ws = torch.empty(need, dtype=torch.uint8, device=o.device)
    parts = None
    if return_order:
        parts = dict(order=torch.empty(rows, n, dtype=torch.int32,
                                       device=o.device),
                     pooled_oracle=torch.empty(rows, n, dtype=torch.float32,
                                               device=o.device),
                     pooled_pred=torch.empty(rows, n, dtype=torch.float32,
                                             device=o.device))
    o2, p2 = o.view(rows, h, w), p.view(rows, h, w)
    stream = torch.cuda.current_stream(o.device).cuda_stream
    for r0 in range(0, rows, chunk):
        r1 = min(rows, r0 + chunk)
        check(L.usl_spars_curve(
            o2[r0:r1].data_ptr(), p2[r0:r1].data_ptr(), r1 - r0, h, w,
            kernel_size, cuts_c, steps, acc.data_ptr(),
            parts['order'][r0:r1].data_ptr() if parts else None,
            parts['pooled_oracle'][r0:r1].data_ptr() if parts else None,
            parts['pooled_pred'][r0:r1].data_ptr() if parts else None,
            ws.data_ptr(), need, stream), 'usl_spars_curve')
    return acc, rows, parts


def finish_curve(row_norm_sum: Tensor, total_rows: int) -> Tensor:
    steps = row_norm_sum.numel()
    out = torch.empty(steps, dtype=torch.float32, device=row_norm_sum.device)
    check(lib().usl_spars_finish(
        row_norm_sum.data_ptr(), steps, total_rows, out.data_ptr(),
        torch.cuda.current_stream(out.device).cuda_stream), 'usl_spars_finish')
    return out


def curve(oracle_error: Tensor, predicted_error: Tensor, kernel_size: int = 11,
          steps: int = 100, device: Device = 'cpu') -> Tensor:
    """Sparsification curve (reference sparsification.py:8-36): mean oracle
    error of the pixels that remain after removing the `step`% most uncertain
    ones, normalised by the full mean, averaged over frames and views.
    Returns a float32 tensor of `steps` values on `device`."""
    acc, rows, _ = curve_sums(oracle_error, predicted_error, kernel_size,
                              steps)
    return finish_curve(acc, rows).to(device)


def random_curve(oracle_error: Tensor, kernel_size: int = 11, steps: int = 100,
                 device: Device = 'cpu') -> Tensor:
    """Curve under a uniformly random ranking (sparsification.py:39-43)."""
    return curve(oracle_error, torch.rand_like(oracle_error), kernel_size,
                 steps, device)


def error(oracle_curve: Tensor, predicted_curve: Tensor) -> Tensor:
    """Sparsification error (sparsification.py:46-49)."""
    return predicted_curve - oracle_curve


def ause(oracle_curve: Tensor, predicted_curve: Tensor) -> Tensor:
    """Area under the sparsification error (sparsification.py:52-57)."""
    if len(oracle_curve) != len(predicted_curve):
        raise Exception('Oracle and Predicted sparsification '
                        'curves have different step sizes.')
    steps = len(oracle_curve)
    if oracle_curve.is_cuda and predicted_curve.is_cuda:
        o = oracle_curve.contiguous().float()
        p = predicted_curve.contiguous().float()
        out = torch.empty((), dtype=torch.float32, device=o.device)
        check(lib().usl_spars_ause(
            o.data_ptr(), p.data_ptr(), steps, out.data_ptr(),
            torch.cuda.current_stream(o.device).cuda_stream), 'usl_spars_ause')
        return out
    # curves already brought to the host (the reference's default
    # device='cpu'): 100 scalars, same canonical order as the kernel
    diff = (predicted_curve.float() - oracle_curve.float()).tolist()
    acc = 0.0
    for d in diff:
        acc += d
    return torch.tensor(acc / steps, dtype=torch.float32)


def aurg(predicted_curve: Tensor, random_curve: Tensor) -> Tensor:
    """Area under the random gain (sparsification.py:60-61)."""
    return ause(predicted_curve, random_curve)
