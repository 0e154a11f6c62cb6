"""TEST INFRASTRUCTURE ONLY -- CPU oracle of the evaluation loop's SSIM metric.

evaluate.py:142-146 calls torchmetrics'
`structural_similarity_index_measure(recon, image, kernel_size=11,
reduction='sum', data_range=1.0)`.  torchmetrics is a third-party dependency
that is NOT vendored in the reference tree (requirements.txt lists it
unpinned, no lock file) and is not installed in this image, so this module
restates the published algorithm of torchmetrics 1.x
(`torchmetrics/functional/image/ssim.py`: `_gaussian`, `_gaussian_kernel_2d`,
`_ssim_update`, `_ssim_compute`) with plain torch ops.  PARITY UNPINNED: there
is no torchmetrics here to generate golden values from; the restatement is
anchored on the reference's call site (its keyword arguments) and on
properties (identical images -> 1 per image, symmetric in its arguments).

Only `tests/` and tools/eval_bench.py's CPU leg may import this module.
"""
import torch
import torch.nn.functional as F
from torch import Tensor


def _gaussian(kernel_size: int, sigma: float, dtype) -> Tensor:
    dist = torch.arange((1 - kernel_size) / 2, (1 + kernel_size) / 2, 1,
                        dtype=dtype)
    gauss = torch.exp(-torch.pow(dist / sigma, 2) / 2)
    return (gauss / gauss.sum()).unsqueeze(0)                 # (1, k)


def ssim_per_image(preds: Tensor, target: Tensor, kernel_size: int = 11,
                   sigma: float = 1.5, data_range: float = 1.0,
                   k1: float = 0.01, k2: float = 0.03) -> Tensor:
    """(B,C,H,W) x 2 -> (B,) SSIM values (gaussian_kernel=True)."""
    c = preds.size(1)
    dtype = preds.dtype
    c1 = (k1 * data_range) ** 2
    c2 = (k2 * data_range) ** 2
    g = _gaussian(kernel_size, sigma, dtype)
    kernel = torch.matmul(g.t(), g).expand(c, 1, kernel_size, kernel_size)
    pad = (kernel_size - 1) // 2
    p = F.pad(preds, (pad, pad, pad, pad), mode='reflect')
    t = F.pad(target, (pad, pad, pad, pad), mode='reflect')
    stack = torch.cat((p, t, p * p, t * t, p * t))            # (5B,C,H+2p,W+2p)
    out = F.conv2d(stack, kernel, groups=c)
    mu_p, mu_t, e_pp, e_tt, e_pt = out.split(preds.size(0))
    mu_pp, mu_tt, mu_pt = mu_p.pow(2), mu_t.pow(2), mu_p * mu_t
    s_pp = torch.clamp(e_pp - mu_pp, min=0.0)
    s_tt = torch.clamp(e_tt - mu_tt, min=0.0)
    s_pt = e_pt - mu_pt
    upper = 2 * s_pt + c2
    lower = s_pp + s_tt + c2
    full = ((2 * mu_pt + c1) * upper) / ((mu_pp + mu_tt + c1) * lower)
    idx = full[..., pad:-pad, pad:-pad]
    return idx.reshape(idx.size(0), -1).mean(-1)


def ssim_sum(preds: Tensor, target: Tensor, **kw) -> Tensor:
    """reduction='sum' (evaluate.py:142-146)."""
    return ssim_per_image(preds, target, **kw).sum()
