"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the sparsification / AUSE path.

Two restatements of `train/sparsification.py` (all citations relative to the
reference tree):

  * `curve_reference_style` follows sparsification.py:8-36 op for op with
    torch on the CPU in fp32 (avg_pool2d -> argsort -> gather -> 100 tail
    means), except that the argsort is made *stable* (`stable=True`), which
    pins the order the reference leaves unspecified on ties to
    (key descending, index ascending).  It is the CPU baseline that is timed
    and the fp32 value the product must match to 1e-6.

  * `pool_exact`, `stable_order`, `curve_canonical` are a numpy restatement
    with every summation order spelled out.  The CUDA kernels implement the
    *same* orders, so pooled keys, the sort permutation, the curve and the
    AUSE are compared bit for bit.

Canonical definitions (shared with `csrc/spars.cu`):
  pool     fp32, window elements added one by one in row-major order starting
           from 0.0f, then one true fp32 division by k*k (what ATen does for
           avg_pool2d; verified against torch in tests/test_oracle.py).
  order    LSD radix order on the key: key descending, ties by ascending
           flat index; -0.0 is treated as +0.0.
  cuts     removed_k = int(k / steps * N) in Python float arithmetic
           (sparsification.py:26-27), computed on the host.
  segment  T_k = sum of the sorted oracle values with rank in
           [removed_k, removed_{k+1}) in fp64: 256 lanes, lane l adds elements
           l, l+256, ... sequentially, then a stride-halving tree 128..1.
  tail     S_99 = T_99, S_k = T_k + S_{k+1} (fp64).
  row      norm_k = (S_k / (N - removed_k)) / (S_0 / N) (fp64).
  curve    curve_k = fp32( (sum over rows in row order of norm_k) / rows ).
  ause     d_k = fp32(pred_k - oracle_k); the d_k are accumulated in fp64 in
           k order, divided by steps, rounded to fp32
           (sparsification.py:46-57).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs
may import this module.
"""
from typing import Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor

LANES = 256


# --------------------------------------------------------------------------
# reference-style (torch, fp32)
# --------------------------------------------------------------------------
def curve_reference_style(oracle_error: Tensor, predicted_error: Tensor,
                          kernel_size: int = 11, steps: int = 100) -> Tensor:
    """sparsification.py:8-36 with a stable argsort."""
    b = predicted_error.size(0)
    o = F.avg_pool2d(oracle_error, kernel_size, stride=1).view(b, 2, -1)
    p = F.avg_pool2d(predicted_error, kernel_size, stride=1).view(b, 2, -1)
    idx = p.argsort(dim=2, descending=True, stable=True)
    o_sorted = o.gather(2, idx)
    o_mean = o.mean(dim=2)
    n = o.size(2)
    out = []
    for step in range(steps):
        removed = int(step / steps * n)
        out.append((o_sorted[:, :, removed:].mean(dim=2) / o_mean).mean())
    return torch.tensor(out)


def ause_reference_style(oracle_curve: Tensor, predicted_curve: Tensor):
    """sparsification.py:46-57."""
    if len(oracle_curve) != len(predicted_curve):
        raise Exception('Oracle and Predicted sparsification '
                        'curves have different step sizes.')
    return (predicted_curve - oracle_curve).sum() / len(oracle_curve)


# --------------------------------------------------------------------------
# canonical (numpy, explicit orders)
# --------------------------------------------------------------------------
def pool_exact(x: np.ndarray, k: int = 11) -> np.ndarray:
    """(..., H, W) fp32 -> (..., H-k+1, W-k+1) fp32, canonical `pool`."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    h, w = x.shape[-2:]
    oh, ow = h - k + 1, w - k + 1
    acc = np.zeros(x.shape[:-2] + (oh, ow), dtype=np.float32)
    for dy in range(k):
        for dx in range(k):
            acc = (acc + x[..., dy:dy + oh, dx:dx + ow]).astype(np.float32)
    return (acc / np.float32(k * k)).astype(np.float32)


def stable_order(keys: np.ndarray) -> np.ndarray:
    """(rows, N) fp32 -> (rows, N) int64 permutation, canonical `order`."""
    keys = np.asarray(keys, dtype=np.float32) + np.float32(0.0)
    return np.argsort(-keys, axis=-1, kind='stable')


def cut_points(n: int, steps: int = 100) -> np.ndarray:
    """sparsification.py:26-27, plus the end sentinel n."""
    return np.array([int(s / steps * n) for s in range(steps)] + [n],
                    dtype=np.int64)


def _segment_sum(seg: np.ndarray) -> np.float64:
    m = seg.shape[0]
    pad = (-m) % LANES
    a = np.concatenate([seg.astype(np.float64), np.zeros(pad)]) \
        .reshape(-1, LANES)
    lanes = np.zeros(LANES, dtype=np.float64)
    for r in range(a.shape[0]):          # sequential per lane
        lanes = lanes + a[r]
    s = LANES // 2
    while s >= 1:                        # stride-halving tree
        lanes[:s] = lanes[:s] + lanes[s:2 * s]
        s //= 2
    return lanes[0]


def row_tail_sums(sorted_vals: np.ndarray, cuts: np.ndarray) -> np.ndarray:
    """(N,) fp32 sorted oracle values -> (steps,) fp64 canonical `tail`."""
    steps = len(cuts) - 1
    seg = np.array([_segment_sum(sorted_vals[cuts[k]:cuts[k + 1]])
                    for k in range(steps)], dtype=np.float64)
    tail = np.zeros(steps, dtype=np.float64)
    run = np.float64(0.0)
    for k in range(steps - 1, -1, -1):
        run = seg[k] + run if k < steps - 1 else seg[k]
        tail[k] = run
    return tail


def curve_canonical(oracle_error: np.ndarray, predicted_error: np.ndarray,
                    kernel_size: int = 11, steps: int = 100,
                    return_parts: bool = False):
    """Canonical curve for (B, 2, H, W) fp32 maps -> (steps,) fp32."""
    o = pool_exact(oracle_error, kernel_size)
    p = pool_exact(predicted_error, kernel_size)
    rows = o.shape[0] * o.shape[1]
    o = o.reshape(rows, -1)
    p = p.reshape(rows, -1)
    n = o.shape[1]
    cuts = cut_points(n, steps)
    order = stable_order(p)
    acc = np.zeros(steps, dtype=np.float64)
    for r in range(rows):
        tail = row_tail_sums(o[r][order[r]], cuts)
        norm = (tail / (n - cuts[:-1]).astype(np.float64)) \
            / (tail[0] / np.float64(n))
        acc = acc + norm
    curve = (acc / np.float64(rows)).astype(np.float32)
    if return_parts:
        return curve, dict(order=order, pooled_oracle=o, pooled_pred=p,
                           row_norm_sum=acc, cuts=cuts)
    return curve


def ause_canonical(oracle_curve: np.ndarray,
                   predicted_curve: np.ndarray) -> np.float32:
    if len(oracle_curve) != len(predicted_curve):
        raise Exception('Oracle and Predicted sparsification '
                        'curves have different step sizes.')
    acc = np.float64(0.0)
    for k in range(len(oracle_curve)):
        acc = acc + np.float64(np.float32(predicted_curve[k])
                               - np.float32(oracle_curve[k]))
    return np.float32(acc / np.float64(len(oracle_curve)))


def synthetic_maps(frames: int, h: int, w: int, seed: int = 0,
                   ties: bool = False) -> Tuple[Tensor, Tensor]:
    """SURVEY.md section 8d, config 5 inputs."""
    g = torch.Generator().manual_seed(seed)
    err = torch.rand(frames, 2, h, w, generator=g)
    unc = torch.clamp(err + 0.2 * torch.rand(frames, 2, h, w, generator=g),
                      0, 1)
    if ties:
        err = torch.round(err * 255) / 255
        unc = torch.round(unc * 255) / 255
    return err, unc
