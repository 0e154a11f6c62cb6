"""Evaluation front half -- the per-batch body of the reference's
`evaluate_model` (train/evaluate.py:136-160) on the GPU.

The drop-in modules (`utils.reconstruct_*_image`, `WeightedSSIMLoss(alpha=1)
.image_error`, `sparsification.curve / random_curve / ause / aurg`) run that
loop unchanged.  `evaluate_batch` is the same sequence as one call, fused where
the data allows: ONE launch of the column kernels warps both views and writes
the reconstructions AND the alpha = 1 error map (the reference's identity
`F.interpolate` of the error to its own size is skipped), the Gaussian SSIM
metric reads the reconstructions once, and the three sparsification curves
share the pooled error.
"""
import ctypes as C
from typing import Dict, Optional

import torch
from torch import Tensor

from .. import functional as K
from .._lib import TERM_REPROJ, check, lib
from ..functional import FusedLoss, LossSettings, ScaleSpec
from . import sparsification as spars
from .utils import Device


def ssim(preds: Tensor, target: Tensor, kernel_size: int = 11,
         sigma: float = 1.5, reduction: Optional[str] = 'sum',
         data_range: float = 1.0, k1: float = 0.01, k2: float = 0.03) -> Tensor:
    """torchmetrics' `structural_similarity_index_measure` with a Gaussian
    window, as evaluate.py:142-146 calls it.  reduction: 'sum' |
    'elementwise_mean' | None / 'none' (per image)."""
    K.require_cuda_f32(preds, 'preds')
    K.require_cuda_f32(target, 'target')
    if preds.shape != target.shape or preds.dim() != 4:
        raise ValueError('preds and target must both be (B,C,H,W)')
    p, t = K.planes(preds), K.planes(target)
    b, c, h, w = p.shape
    L = lib()
    need = L.usl_ssim_workspace_bytes(b, c, h, w, kernel_size)
    if need == 0:
        raise ValueError('unsupported SSIM shape / kernel size')
    ws = torch.empty(need, dtype=torch.uint8, device=p.device)
    out = torch.empty(b, dtype=torch.float32, device=p.device)
    check(L.usl_ssim_gauss(p.data_ptr(), p.stride(0), p.stride(1), t.data_ptr(),
                           t.stride(0), t.stride(1), b, c, h, w, kernel_size,
                           float(sigma), float(data_range), float(k1), float(k2),
                           out.data_ptr(), ws.data_ptr(), need,
                           torch.cuda.current_stream(p.device).cuda_stream),
          'usl_ssim_gauss')
    if reduction == 'sum':
        return out.sum()
    if reduction == 'elementwise_mean':
        return out.mean()
    if reduction in (None, 'none'):
        return out
    raise ValueError(f'unknown reduction {reduction!r}')


def reconstruct_and_error(images: Tensor, prediction: Tensor):
    """evaluate.py:139-153 in one launch: both reconstructions (B,6,H,W) and
    the alpha = 1 image error (B,2,H,W) of a stereo pair and a prediction."""
    K.require_cuda_f32(images, 'images')
    K.require_cuda_f32(prediction, 'prediction')
    b, _, h, w = images.shape
    with torch.no_grad():
        spec = ScaleSpec(terms=TERM_REPROJ,
                         coefs=(1.0 / (b * h * w), 0, 0, 0, 0, 0), images=0,
                         disp=1, disp_ch=0, want_err=True, want_recon=True)
        out = FusedLoss.apply(LossSettings(alpha=1.0), [spec], None, images,
                              prediction)
    return out[4], out[3]


def evaluate_batch(left: Tensor, right: Tensor, prediction: Tensor,
                   kernel_size: int = 11, device: Device = 'cpu',
                   random_error: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """One iteration of the reference's evaluation loop (evaluate.py:136-160):
    summed SSIM of either reconstruction, AUSE and AURG of the batch."""
    images = torch.cat([left, right], dim=1)
    uncertainty = prediction[:, 2:4]
    recon, error = reconstruct_and_error(images, prediction)
    left_ssim = ssim(recon[:, 0:3], left, kernel_size=kernel_size,
                     reduction='sum', data_range=1.0)
    right_ssim = ssim(recon[:, 3:6], right, kernel_size=kernel_size,
                      reduction='sum', data_range=1.0)
    oracle_spars = spars.curve(error, error, device=device)
    pred_spars = spars.curve(error, uncertainty, device=device)
    if random_error is None:
        random_error = torch.rand_like(error)
    random_spars = spars.curve(error, random_error, device=device)
    return dict(left_ssim=left_ssim, right_ssim=right_ssim,
                ause=spars.ause(oracle_spars, pred_spars),
                aurg=spars.aurg(pred_spars, random_spars),
                recon=recon, error=error, oracle_curve=oracle_spars,
                pred_curve=pred_spars)
