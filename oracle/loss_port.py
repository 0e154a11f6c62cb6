"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the loss hot path.

A restatement ("port") of the reference's multi-scale stereo loss in plain
functional PyTorch on CPU tensors.  It follows the reference's own op
sequence (same ATen library calls in the same order: `F.interpolate`,
`F.grid_sample`, `avg_pool2d`, replicate `F.pad`), so that

  * its results are the reference's results (pinned by the fixtures in
    `tests/golden/`, generated from the real reference by
    `oracle/make_golden.py`), in fp32 or -- by feeding it float64 tensors --
    in fp64 as the tolerance arbiter, with autograd providing gradients;
  * its CPU time is representative of the reference's CPU path, which is
    what `bench.py`'s `cpu_baseline` / `--impl reference` legs report.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs
may import this module.  The product (`uncertainty_model_b200`) never does.

All `file:line` citations are relative to the reference tree.
"""
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

LOSS_TYPES = ('l1', 'bayesian', 'log_bayesian')


# --------------------------------------------------------------------------
# train/utils.py
# --------------------------------------------------------------------------
def pyramid(x: Tensor, scales: int = 4) -> List[Tensor]:
    """utils.py:27-50 -- every level is resampled from the full-res input."""
    h, w = x.shape[-2:]
    return [F.interpolate(x, size=(h // 2 ** i, w // 2 ** i),
                          mode='bilinear', align_corners=True)
            for i in range(scales)]


def warp(shift: Tensor, source: Tensor) -> Tensor:
    """utils.py:65-97 -- horizontal shift in normalised [0, 1] units.

    The base grid is linspace(0, 1, n) while grid_sample runs with its default
    align_corners=False; this mismatch is part of the reference's behaviour.
    """
    b, _, h, w = source.shape
    xs = torch.linspace(0, 1, w).repeat(b, h, 1).type_as(source)
    ys = torch.linspace(0, 1, h).repeat(b, w, 1).transpose(1, 2) \
        .type_as(source)
    grid = torch.stack((xs + shift.squeeze(1), ys), dim=3)
    grid = (2 * grid) - 1
    return F.grid_sample(source, grid, mode='bilinear', padding_mode='zeros',
                         align_corners=False)


def warp_to_left(disp_left: Tensor, right: Tensor) -> Tensor:
    """utils.py:100-103."""
    return warp(-disp_left, right)


def warp_to_right(disp_right: Tensor, left: Tensor) -> Tensor:
    """utils.py:106-109."""
    return warp(disp_right, left)


def recon_pyramid(preds: Sequence[Tensor],
                  images: Sequence[Tensor]) -> List[Tensor]:
    """utils.py:112-135."""
    out = []
    for p, im in zip(preds, images):
        out.append(torch.cat((warp_to_left(p[:, 0:1], im[:, 3:6]),
                              warp_to_right(p[:, 1:2], im[:, 0:3])), dim=1))
    return out


# --------------------------------------------------------------------------
# train/loss.py
# --------------------------------------------------------------------------
def _pool3(x: Tensor) -> Tensor:
    return F.avg_pool2d(x, kernel_size=3, stride=1)


def ssim_map(x: Tensor, y: Tensor, c1: float, c2: float) -> Tensor:
    """loss.py:43-74 (c1, c2 are the already squared k1, k2 of loss.py:31-32)."""
    mx, my = _pool3(x), _pool3(y)
    mxx, myy, mxy = mx * mx, my * my, mx * my
    vx = _pool3(x * x) - mxx
    vy = _pool3(y * y) - myy
    vxy = _pool3(x * y) - mxy
    return (((2 * mxy) + c1) * ((2 * vxy) + c2)) \
        / ((mxx + myy + c1) * (vx + vy + c2))


def image_error(images: Tensor, recon: Tensor, alpha: float = 0.85,
                k1: float = 0.01, k2: float = 0.03) -> Tensor:
    """loss.py:96-131 -> (B, 2, h, w) per-view photometric error."""
    h, w = images.shape[-2:]
    c1, c2 = k1 ** 2, k2 ** 2
    l1 = (images - recon).abs()
    ds = torch.cat([torch.clamp((1 - ssim_map(images[:, s], recon[:, s],
                                              c1, c2)) / 2, 0, 1)
                    for s in (slice(0, 3), slice(3, 6))], dim=1)
    ds = F.interpolate(ds, size=(h, w), mode='bilinear', align_corners=True)
    total = (alpha * ds) + ((1 - alpha) * l1)
    return torch.cat((total[:, 0:3].mean(dim=1, keepdim=True),
                      total[:, 3:6].mean(dim=1, keepdim=True)), dim=1)


def consistency(a: Tensor, b: Optional[Tensor] = None) -> Tensor:
    """loss.py:167-188 -- `a` supplies the shift, `b` (default `a`) is sampled."""
    b = a if b is None else b
    left = (a[:, 0:1] - warp_to_left(a[:, 0:1], b[:, 1:2])).abs().mean()
    right = (a[:, 1:2] - warp_to_right(a[:, 1:2], b[:, 0:1])).abs().mean()
    return left + right


def _dx(x: Tensor) -> Tensor:
    x = F.pad(x, (0, 1, 0, 0), mode='replicate')          # loss.py:211-212
    return x[..., :-1] - x[..., 1:]


def _dy(x: Tensor) -> Tensor:
    x = F.pad(x, (0, 0, 0, 1), mode='replicate')          # loss.py:217-218
    return x[..., :-1, :] - x[..., 1:, :]


def smoothness(disp: Tensor, images: Tensor) -> Tensor:
    """loss.py:224-264."""
    total = 0
    for d, im in ((disp[:, 0:1], images[:, 0:3]),
                  (disp[:, 1:2], images[:, 3:6])):
        wx = torch.exp(-_dx(im).abs().mean(dim=1, keepdim=True))
        wy = torch.exp(-_dy(im).abs().mean(dim=1, keepdim=True))
        total = total + (_dx(d) * wx).abs() + (_dy(d) * wy).abs()
    return total.mean()


def uncertainty_loss(pred: Tensor, images: Tensor, error: Tensor,
                     loss_type: str = 'l1', smoothness_weight: float = 1.0,
                     consistency_weight: float = 1.0,
                     pooling: bool = False) -> Tensor:
    """loss.py:340-434."""
    if loss_type not in LOSS_TYPES:
        raise ValueError('Loss must be either "l1", "bayesian" '
                         'or "log_bayesian".')
    error = error.detach().clone()                         # loss.py:418
    if pooling:                                            # loss.py:420-422
        pred, images, error = _pool3(pred), _pool3(images), _pool3(error)
    disp, unc = pred[:, 0:2], pred[:, 2:4]
    if loss_type == 'l1':
        loss = (unc - error).abs().mean()
    elif loss_type == 'bayesian':
        loss = ((error / unc) + torch.log(unc)).mean()
    else:
        loss = ((error / torch.exp(-unc)) + unc).mean() / 2
    if smoothness_weight > 0:
        loss = loss + smoothness(unc, images) * smoothness_weight
    if consistency_weight > 0:
        loss = loss + consistency(unc, disp) * consistency_weight
    return loss


DEFAULT_LOSS_CONFIG = dict(
    wssim_weight=1.0, consistency_weight=1.0, smoothness_weight=1.0,
    adversarial_weight=0.85, predictive_error_weight=1.0,
    perceptual_weight=0.05, wssim_alpha=0.85, perceptual_start=5,
    adversarial_loss_type='mse', error_loss_config=None)


def total_loss(images: Sequence[Tensor], preds: Sequence[Tensor],
               recons: Sequence[Tensor], config: Optional[dict] = None,
               return_terms: bool = False):
    """loss.py:512-568 without a discriminator -> (disp_loss, error_loss)."""
    cfg = dict(DEFAULT_LOSS_CONFIG)
    cfg.update(config or {})
    err_cfg = cfg['error_loss_config'] or {}
    reproj = cons = smooth = err = 0
    errors = []
    for i, (im, p, rc) in enumerate(zip(images, preds, recons)):
        e = image_error(im, rc, cfg['wssim_alpha'])
        errors.append(e)
        reproj = reproj + (e[:, 0:1] + e[:, 1:2]).mean()   # loss.py:146-151
        cons = cons + consistency(p[:, 0:2])
        smooth = smooth + smoothness(p[:, 0:2], im) / (2 ** i)
        err = err + uncertainty_loss(p, im, e, **err_cfg)
    disp_loss = reproj * cfg['wssim_weight'] \
        + cons * cfg['consistency_weight'] \
        + smooth * cfg['smoothness_weight']
    error_loss = err * cfg['predictive_error_weight']
    if return_terms:
        return disp_loss, error_loss, dict(reproj=reproj, cons=cons,
                                           smooth=smooth, err=err,
                                           errors=errors)
    return disp_loss, error_loss


def step(stereo: Tensor, preds: Sequence[Tensor],
         config: Optional[dict] = None) -> Tuple[Tensor, Tensor, List[Tensor]]:
    """One pass of the hot path as train.py:117-128 runs it: pyramid, warp,
    loss forward, backward.  Returns (disp_loss, error_loss, grad(preds))."""
    preds = [p.detach().clone().requires_grad_(True) for p in preds]
    pyr = pyramid(stereo, len(preds))
    rec = recon_pyramid(preds, pyr)
    dl, el = total_loss(pyr, preds, rec, config)
    (dl + el).backward()
    return dl.detach(), el.detach(), [p.grad for p in preds]


def step_detailed(stereo: Tensor, preds: Sequence[Tensor],
                  config: Optional[dict] = None) -> dict:
    """`step` plus the intermediates a parity test needs to say WHERE the
    gradient is one-sided (oracle/kinks.py): the pyramid, the reconstructions
    and the per-scale error maps."""
    preds = [p.detach().clone().requires_grad_(True) for p in preds]
    pyr = pyramid(stereo, len(preds))
    rec = recon_pyramid(preds, pyr)
    dl, el, terms = total_loss(pyr, preds, rec, config, return_terms=True)
    (dl + el).backward()
    return dict(disp_loss=dl.detach(), error_loss=el.detach(),
                grads=[p.grad for p in preds], pyramid=pyr,
                preds=[p.detach() for p in preds],
                recons=[r.detach() for r in rec],
                errors=[e.detach() for e in terms['errors']])
