"""TEST INFRASTRUCTURE ONLY -- where the loss is not differentiable.

The loss of train/loss.py is piecewise smooth in the predictions: its
gradient jumps wherever

  K1  the sampling coordinate of a warp crosses an integer (the two bilinear
      taps change, utils.py:93-97): d(recon)/d(shift) is one-sided;
  K2  an absolute value changes sign: |I - recon| (loss.py:92-94),
      |a - warp(b)| (loss.py:182-186), |d[j] - d[j+1]| (loss.py:241-246),
      |u - E| (loss.py:401-403);
  K3  the clamp of the dissimilarity binds (loss.py:89-90).

Within fp32 rounding of such a point an fp32 evaluation (the reference's own,
the CUDA kernels') and the fp64 arbiter may sit on different sides, and the
affected gradient ELEMENTS then differ by O(their size) -- the one-sided
derivatives are both "right".  `kink_masks` marks exactly those elements, from
the fp64 intermediates and explicit margins, so that a parity test can compare
every other element strictly and still account for each excluded one.

Margins.  With xs = linspace(0,1,w), ix = ((2 (xs + shift) - 1) + 1) * w/2 - 1/2
is evaluated in fp32 (utils.py:80-94, then ATen's unnormalise; the kernels use
the same sequence with fused multiply-adds): three roundings of O(1) values
(<= 1.2e-7 each) are amplified by w/2 and the result is rounded once more
(<= 3e-5 at 512), i.e. |ix32 - ix64| <= IX_EPS_PER_W * w for any fp32
evaluation order (the reference's own stays within a fifth of it, checked in
tests/test_oracle.py); likewise iy with h.  A
warped value inherits |d out/d ix| * err(ix) + |d out/d iy| * err(iy) from its
weights, plus its own rounding; sign margins are built from that per element.
"""
from typing import Dict, List, Sequence

import torch
from torch import Tensor

IX_EPS_PER_W = 2.5e-7      # fp32 error bound of a sampling coordinate, per column
ABS_EPS = 4e-7             # fp32 rounding of an O(1) value (a few ulp)
CLAMP_EPS = 1e-5           # dssim within this of 0 or 1 (ssim is a quotient)


def explicit_warp(shift: Tensor, src: Tensor) -> Dict[str, Tensor]:
    """utils.py:65-97 written out tap by tap (same values as
    oracle.loss_port.warp, checked in tests): returns the warped image and
    the taps / slopes the masks need."""
    b, c, h, w = src.shape
    dt = src.dtype
    xs = torch.linspace(0, 1, w).to(dt).view(1, 1, w)
    ys = torch.linspace(0, 1, h).to(dt).view(1, h, 1)
    gx = 2 * (xs + shift.squeeze(1)) - 1
    gy = (2 * ys - 1).expand(b, h, w)
    ix = ((gx + 1) * w - 1) / 2
    iy = ((gy + 1) * h - 1) / 2
    x0 = torch.floor(ix)
    y0 = torch.floor(iy)
    fx, fy = ix - x0, iy - y0
    x0, y0 = x0.long(), y0.long()

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
        idx = (yy.clamp(0, h - 1) * w + xx.clamp(0, w - 1)).view(b, 1, -1) \
            .expand(b, c, -1)
        v = src.reshape(b, c, -1).gather(2, idx).view(b, c, h, w)
        return v * ok.view(b, 1, h, w).to(dt)

    fxb, fyb = fx.unsqueeze(1), fy.unsqueeze(1)
    a00, a01 = tap(y0, x0), tap(y0, x0 + 1)
    a10, a11 = tap(y0 + 1, x0), tap(y0 + 1, x0 + 1)
    t0 = (1 - fyb) * a00 + fyb * a10                            # column x0
    t1 = (1 - fyb) * a01 + fyb * a11
    out = (1 - fxb) * t0 + fxb * t1
    # d out / d ix and d out / d iy: what an error of the coordinates costs
    slope_y = ((1 - fxb) * a10 + fxb * a11) - ((1 - fxb) * a00 + fxb * a01)
    return dict(out=out, ix=ix, x0=x0, y0=y0, slope=t1 - t0, slope_y=slope_y)


def _margin(wp: Dict[str, Tensor], eps_ix: float, eps_iy: float) -> Tensor:
    """fp32 error bound of a warped value."""
    return eps_ix * wp['slope'].abs() + eps_iy * wp['slope_y'].abs() + ABS_EPS


def _near_integer(ix: Tensor, eps: float) -> Tensor:
    return (ix - torch.round(ix)).abs() <= eps


def _mark_taps(mask_plane: Tensor, sel: Tensor, y0: Tensor, x0: Tensor) -> None:
    """mask_plane (B,h,w) |= the four source taps of every selected pixel."""
    b, h, w = mask_plane.shape
    if not bool(sel.any()):
        return
    bi = torch.arange(b).view(b, 1, 1).expand(b, h, w)[sel]
    yy, xx = y0[sel], x0[sel]
    for dy in (0, 1):
        for dx in (0, 1):
            y, x = yy + dy, xx + dx
            ok = (y >= 0) & (y < h) & (x >= 0) & (x < w)
            mask_plane[bi[ok], y[ok], x[ok]] = True


def warp_kinks(shift: Tensor, src: Tensor) -> Tensor:
    """K1 alone: (B,h,w) pixels whose sampling column sits on an integer."""
    return _near_integer(explicit_warp(shift, src)['ix'],
                         IX_EPS_PER_W * src.shape[-1])


def consistency_kinks(a: Tensor, b: Tensor):
    """ConsistencyLoss(a, b) (loss.py:167-188): `a` (B,2,h,w) supplies the
    shift and is compared with the warped `b`.  Returns the masks of the
    gradient w.r.t. a (K1 and the sign of a - warp(b)) and w.r.t. b (the four
    taps a pixel with an ambiguous sign scatters to)."""
    bsz, _, h, w = a.shape
    ma = torch.zeros(bsz, 2, h, w, dtype=torch.bool)
    mb = torch.zeros(bsz, 2, h, w, dtype=torch.bool)
    eps_ix, eps_iy = IX_EPS_PER_W * w, IX_EPS_PER_W * h
    for v, sign in ((0, -1.0), (1, 1.0)):
        av = a[:, v:v + 1]
        wp = explicit_warp(sign * av, b[:, 1 - v:2 - v])
        sel = (av - wp['out']).abs().squeeze(1) <= \
            _margin(wp, eps_ix, eps_iy).squeeze(1)
        ma[:, v] |= _near_integer(wp['ix'], eps_ix) | sel
        _mark_taps(mb[:, 1 - v], sel, wp['y0'], wp['x0'])
    return ma, mb


def smoothness_kinks(a: Tensor) -> Tensor:
    """SmoothnessLoss (loss.py:241-246): differences of two inputs are exact
    in fp32, only exact ties are ambiguous (both ends of the edge)."""
    m = torch.zeros(a.shape, dtype=torch.bool)
    tx = (a[..., :, :-1] - a[..., :, 1:]).abs() <= 0.25 * ABS_EPS
    ty = (a[..., :-1, :] - a[..., 1:, :]).abs() <= 0.25 * ABS_EPS
    m[..., :, :-1] |= tx
    m[..., :, 1:] |= tx
    m[..., :-1, :] |= ty
    m[..., 1:, :] |= ty
    return m


def _disp_term_kinks(im: Tensor, p: Tensor, alpha: float,
                     image_eps: float = 0.0) -> Tensor:
    """Kinks of the disparity terms of one scale (reprojection, consistency of
    the disparity, its smoothness): (B,4,h,w), uncertainty channels untouched.
    `image_eps`: how far an fp32 pyramid level is from the fp64 one (levels
    below the first are interpolated: on white-noise images the fp32 rounding
    of the interpolation weights moves a pixel by ~1e-6)."""
    b, _, h, w = p.shape
    m = torch.zeros(b, 4, h, w, dtype=torch.bool)
    eps_ix, eps_iy = IX_EPS_PER_W * w, IX_EPS_PER_W * h
    for v, sign in ((0, -1.0), (1, 1.0)):
        opp_img = im[:, 3:6] if v == 0 else im[:, 0:3]
        own_img = im[:, 0:3] if v == 0 else im[:, 3:6]
        wi = explicit_warp(sign * p[:, v:v + 1], opp_img)
        # K1: taps of the image warp driven by d
        m[:, v] |= _near_integer(wi['ix'], eps_ix)
        # K2: |I - recon| per channel (weight 1 - alpha)
        if alpha < 1.0:
            # (the image and the four pixels the reconstruction blends)
            m[:, v] |= ((own_img - wi['out']).abs() <=
                        _margin(wi, eps_ix, eps_iy) + 2.0 * image_eps).any(dim=1)
    ma, mb = consistency_kinks(p[:, 0:2], p[:, 0:2])
    m[:, 0:2] |= ma | mb | smoothness_kinks(p[:, 0:2])
    return m


def error_term_kinks(p: Tensor, e: Tensor, loss_type: str, u_smooth: bool,
                     u_cons: bool) -> Tensor:
    """Kinks of ReprojectionErrorLoss (loss.py:389-434) on the maps it runs on
    (3x3-pooled ones when `pooling` is set): (B,4,h,w)."""
    b, _, h, w = p.shape
    m = torch.zeros(b, 4, h, w, dtype=torch.bool)
    if u_cons:
        ma, mb = consistency_kinks(p[:, 2:4], p[:, 0:2])
        m[:, 2:4] |= ma
        m[:, 0:2] |= mb
    if u_smooth:
        m[:, 2:4] |= smoothness_kinks(p[:, 2:4])
    if loss_type == 'l1':
        m[:, 2:4] |= (p[:, 2:4] - e).abs() <= 8 * ABS_EPS
    return m


def dilate3(mp: Tensor) -> Tensor:
    """A kink at pooled position q touches the 3x3 inputs under it
    (loss.py:420-422): (.., h-2, w-2) -> (.., h, w)."""
    import torch.nn.functional as F
    return F.max_pool2d(F.pad(mp.double(), (2, 2, 2, 2)), 3, 1) > 0


def kink_masks(pyramid: Sequence[Tensor], preds: Sequence[Tensor],
               config: dict, errors: Sequence[Tensor],
               image_eps: Sequence[float] = ()) -> List[Tensor]:
    """Boolean (B,4,h,w) per scale: True where the gradient element of
    prediction channel {d_L, d_R, u_L, u_R} may legitimately differ between an
    fp32 and an fp64 evaluation.  All inputs fp64 (the oracle's pyramid, the
    predictions, its per-scale error maps E of loss.py:126-131); `image_eps`:
    per scale, max |fp32 pyramid - fp64 pyramid| (0 where omitted)."""
    import torch.nn.functional as F
    err_cfg = config.get('error_loss_config') or {}
    loss_type = err_cfg.get('loss_type', 'l1')
    u_smooth = err_cfg.get('smoothness_weight', 1.0) > 0
    u_cons = err_cfg.get('consistency_weight', 1.0) > 0
    pooling = bool(err_cfg.get('pooling', False))
    alpha = config.get('wssim_alpha', 0.85)
    masks = []
    for i, (im, p, e) in enumerate(zip(pyramid, preds, errors)):
        m = _disp_term_kinks(im, p, alpha,
                             image_eps[i] if i < len(image_eps) else 0.0)
        if pooling:
            # loss.py:420-422: the error terms see 3x3 means; a kink at pooled
            # position q touches the 3x3 inputs under it
            m |= dilate3(error_term_kinks(F.avg_pool2d(p, 3, 1),
                                          F.avg_pool2d(e, 3, 1), loss_type,
                                          u_smooth, u_cons))
        else:
            m |= error_term_kinks(p, e, loss_type, u_smooth, u_cons)
        masks.append(m)
    return masks


def clamp_kinks(pyramid: Sequence[Tensor], recons: Sequence[Tensor]) -> int:
    """K3: number of 3x3 windows whose dissimilarity sits within CLAMP_EPS of
    the clamp bounds (0 on every input the tests use; the tests assert it)."""
    from oracle.loss_port import ssim_map
    n = 0
    for im, rc in zip(pyramid, recons):
        for s in (slice(0, 3), slice(3, 6)):
            raw = (1 - ssim_map(im[:, s], rc[:, s], 1e-4, 9e-4)) / 2
            n += int(((raw.abs() <= CLAMP_EPS) |
                      ((raw - 1).abs() <= CLAMP_EPS)).sum())
    return n
