"""Host side of the hot path: tensors -> C ABI structs -> libusl launches,
wrapped as `torch.autograd.Function`s.

PyTorch is plumbing here (device memory, streams, autograd bookkeeping); all
arithmetic happens in the CUDA kernels behind `_lib.lib()`.  Inputs must be
CUDA fp32 tensors: there is no CPU fallback.
"""
import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import (LOSS_TYPES, TERM_CONS_D, TERM_CONS_U, TERM_REPROJ,
                   TERM_SMOOTH_D, TERM_SMOOTH_U, TERM_UNC, USL_NUM_TERMS,
                   UslLossConfig, UslLossScale, check, lib)


# --------------------------------------------------------------------------
# tensor plumbing
# --------------------------------------------------------------------------
def _stream(t: Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def require_cuda_f32(t: Tensor, name: str) -> None:
    if not isinstance(t, Tensor):
        raise TypeError(f'{name} must be a torch.Tensor')
    if t.dtype != torch.float32:
        raise TypeError(f'{name} must be float32, got {t.dtype}')
    if not t.is_cuda:
        raise ValueError(f'{name} must be a CUDA tensor: the B200 loss path '
                         'has no CPU fallback')


def planes(t: Tensor) -> Tensor:
    """A view/copy of a (B,C,h,w) tensor whose h*w planes are contiguous."""
    if t.dim() != 4:
        raise ValueError(f'expected a 4-d tensor, got shape {tuple(t.shape)}')
    h, w = t.shape[-2:]
    if t.stride(3) == 1 and t.stride(2) == w:
        return t
    return t.contiguous()


class ChannelPair:
    """Two consecutive channels of a (B,C,h,w) tensor, as the C ABI sees them
    (base pointer, batch / channel strides) -- without building a view."""
    __slots__ = ('ptr', 'bs', 'cs')

    def __init__(self, t: Tensor, ch: int) -> None:
        self.bs, self.cs = t.stride(0), t.stride(1)
        self.ptr = t.data_ptr() + 4 * ch * self.cs


def _ptr(t) -> Optional[int]:
    if t is None:
        return None
    return t.ptr if isinstance(t, ChannelPair) else t.data_ptr()


def _strides(t) -> Tuple[int, int]:
    if t is None:
        return 0, 0
    if isinstance(t, ChannelPair):
        return t.bs, t.cs
    return t.stride(0), t.stride(1)


# --------------------------------------------------------------------------
# loss configuration -> per-scale term masks and coefficients
# --------------------------------------------------------------------------
@dataclass(frozen=True)
class LossSettings:
    """The scalars of `config.yml`'s `loss:` block that the kernels need
    (reference: train/loss.py:438-447 and 357-360)."""
    wssim_weight: float = 1.0
    consistency_weight: float = 1.0
    smoothness_weight: float = 1.0
    predictive_error_weight: float = 1.0
    alpha: float = 0.85
    c1: float = 0.01 ** 2
    c2: float = 0.03 ** 2
    loss_type: str = 'l1'
    err_smoothness_weight: float = 1.0
    err_consistency_weight: float = 1.0

    def terms(self) -> int:
        t = TERM_REPROJ | TERM_CONS_D | TERM_SMOOTH_D | TERM_UNC
        if self.err_smoothness_weight > 0:      # loss.py:428
            t |= TERM_SMOOTH_U
        if self.err_consistency_weight > 0:     # loss.py:430
            t |= TERM_CONS_U
        return t

    def coefs(self, scale_index: int, n_pixels: int,
              n_unc: Optional[int] = None) -> List[float]:
        """coef[k] * (raw sum k) = contribution of term k at this scale.

        n_pixels = B*h*w of the (global) batch: every mean of the reference
        (loss.py:151,185-186,264,393-403) is a sum divided by it (or by twice
        it for the two-channel uncertainty mean).  n_unc is the pixel count of
        the error terms when they run on pooled maps."""
        n = float(n_pixels)
        nu = float(n_unc if n_unc is not None else n_pixels)
        pe = self.predictive_error_weight
        half = 0.5 if self.loss_type == 'log_bayesian' else 1.0
        return [self.wssim_weight / n,
                self.consistency_weight / n,
                self.smoothness_weight / (n * 2 ** scale_index),
                pe * half / (2.0 * nu),
                pe * self.err_smoothness_weight / nu,
                pe * self.err_consistency_weight / nu]


def make_config(terms: int, settings: LossSettings,
                coefs: Sequence[float]) -> UslLossConfig:
    cfg = UslLossConfig()
    cfg.terms = terms
    cfg.loss_type = LOSS_TYPES[settings.loss_type]
    cfg.alpha, cfg.c1, cfg.c2 = settings.alpha, settings.c1, settings.c2
    for k in range(USL_NUM_TERMS):
        cfg.coef[k] = coefs[k]
    return cfg


def make_scale(images: Optional[Tensor], disp: Optional[Tensor],
               unc: Optional[Tensor], *, shape: Tuple[int, int, int],
               recon_in: Optional[Tensor] = None,
               err_in: Optional[Tensor] = None,
               recon_out: Optional[Tensor] = None,
               err_out: Optional[Tensor] = None,
               grad_recon_in: Optional[Tensor] = None,
               grad_disp: Optional[Tensor] = None,
               grad_unc: Optional[Tensor] = None,
               grad_recon_out: Optional[Tensor] = None,
               scatter_ws: Optional[Tensor] = None,
               flags: int = 0) -> UslLossScale:
    s = UslLossScale()
    s.B, s.h, s.w = shape
    s.flags = flags
    s.images = _ptr(images); s.img_bs, s.img_cs = _strides(images)
    s.disp = _ptr(disp); s.disp_bs, s.disp_cs = _strides(disp)
    s.unc = _ptr(unc); s.unc_bs, s.unc_cs = _strides(unc)
    s.recon_in = _ptr(recon_in); s.rin_bs, s.rin_cs = _strides(recon_in)
    s.err_in = _ptr(err_in); s.ein_bs, s.ein_cs = _strides(err_in)
    s.recon_out = _ptr(recon_out)
    s.err_out = _ptr(err_out)
    s.grad_recon_in = _ptr(grad_recon_in)
    s.grad_disp = _ptr(grad_disp); s.gd_bs, s.gd_cs = _strides(grad_disp)
    s.grad_unc = _ptr(grad_unc); s.gu_bs, s.gu_cs = _strides(grad_unc)
    s.grad_recon_out = _ptr(grad_recon_out)
    s.scatter_ws = _ptr(scatter_ws)
    return s


def _array(cls, items):
    arr = (cls * len(items))()
    for i, it in enumerate(items):
        arr[i] = it
    return arr


# --------------------------------------------------------------------------
# launches
# --------------------------------------------------------------------------
MODE_FWD, MODE_GRAD = 0, 1
GRAD_SKIP_IF_UNIT, GRAD_NO_SCATTER, GRAD_ONLY_SCATTER = 1, 2, 4
GRAD_DEFER_SCATTER0, GRAD_ONLY_SCATTER0 = 8, 16


def plan_rows(cfg_arr, sc_arr, n: int, mode: int) -> Optional[List[int]]:
    """Row offsets of every scale in the partial-sum buffer for a launch in
    `mode`; None if the one-pass sums+gradient launch does not apply."""
    starts = (C.c_int * (n + 1))()
    rc = lib().usl_loss_plan(cfg_arr, sc_arr, n, mode, starts)
    if rc == _lib.USL_ERR_UNSUPPORTED and mode == MODE_GRAD:
        return None
    check(rc, 'usl_loss_plan')
    return list(starts)


class Reduce:
    """How a sharded batch is put together: the raw term sums are all-reduced
    over `group`; `combine_scale` turns the coefficients the kernels used (and
    scaled the gradients with) into those of the reported loss -- 1 when the
    kernels already normalise by the global batch, 1/world when they normalise
    by the local shard (torch DDP averages parameter gradients itself)."""
    __slots__ = ('group', 'combine_scale')

    def __init__(self, group, combine_scale: float = 1.0) -> None:
        self.group = group
        self.combine_scale = float(combine_scale)


def loss_forward(cfgs: Sequence[UslLossConfig],
                 scales: Sequence[UslLossScale], device,
                 reduce: Optional[Reduce] = None, with_grad: bool = False,
                 arrays=None, starts: Optional[List[int]] = None
                 ) -> Tuple[Tensor, Tensor, Tensor]:
    """One fused launch over all scales.

    Returns (disp_loss, error_loss, sums): two 0-dim fp32 tensors and the
    fp64[n_scales, 6] raw per-term sums (all-reduced over `reduce.group` when
    the batch is sharded over ranks).

    with_grad: use the one-pass launch that also writes the gradients (for
    unit upstream gradients) into the grad_* buffers of `scales`; the caller
    has checked with `plan_rows(..., MODE_GRAD)` that it applies."""
    L = lib()
    n = len(scales)
    cfg_arr, sc_arr = arrays if arrays is not None else (
        _array(UslLossConfig, cfgs), _array(UslLossScale, scales))
    if starts is None:
        starts = plan_rows(cfg_arr, sc_arr, n,
                           MODE_GRAD if with_grad else MODE_FWD)
    partials = torch.empty(starts[-1] * USL_NUM_TERMS, dtype=torch.float32,
                           device=device)
    sums = torch.empty(n, USL_NUM_TERMS, dtype=torch.float64, device=device)
    out_disp = torch.empty((), dtype=torch.float32, device=device)
    out_err = torch.empty((), dtype=torch.float32, device=device)
    reduce_group = reduce.group if reduce is not None else None
    cscale = reduce.combine_scale if reduce is not None else 1.0
    coef = coef_tensor(tuple(tuple(k * cscale for k in c.coef) for c in cfgs),
                       device)
    stream = _stream(partials)
    # With gradients: the reduction of the term sums, their exchange between
    # ranks (sharded batch) and the combination only need the column kernels, so
    # they run on a side stream (-> NCCL stream -> side stream) beside the
    # transposed warps and are joined at the end.
    split = with_grad and reduce_group is not None
    starts_arr = (C.c_int * (n + 1))(*starts)
    if with_grad:
        side = _reduce_stream(partials.device)
        rc = L.usl_loss_grad_sharded(cfg_arr, sc_arr, n, partials.data_ptr(),
                                     starts_arr, sums.data_ptr(),
                                     side.cuda_stream, stream)
        if rc != _lib.USL_ERR_UNSUPPORTED:
            check(rc, 'usl_loss_grad_sharded')
            with torch.cuda.stream(side):
                if reduce_group is not None:
                    torch.distributed.all_reduce(sums, group=reduce_group)
                check(L.usl_loss_combine(sums.data_ptr(), coef.data_ptr(), n,
                                         out_disp.data_ptr(),
                                         out_err.data_ptr(), side.cuda_stream),
                      'usl_loss_combine')
            torch.cuda.current_stream(partials.device).wait_stream(side)
            return out_disp, out_err, sums
        check(L.usl_loss_grad(cfg_arr, sc_arr, n, None, None,
                              partials.data_ptr(),
                              GRAD_DEFER_SCATTER0 if split else 0, stream),
              'usl_loss_grad')
    else:
        check(L.usl_loss_fwd(cfg_arr, sc_arr, n, partials.data_ptr(), stream),
              'usl_loss_fwd')
    check(L.usl_loss_reduce(partials.data_ptr(), starts_arr, n,
                            sums.data_ptr(), stream), 'usl_loss_reduce')
    if split:
        work = torch.distributed.all_reduce(sums, group=reduce_group,
                                            async_op=True)
        check(L.usl_loss_grad(cfg_arr, sc_arr, n, None, None, None,
                              GRAD_ONLY_SCATTER0, stream), 'usl_loss_grad')
        work.wait()
    elif reduce_group is not None:
        torch.distributed.all_reduce(sums, group=reduce_group)
    check(L.usl_loss_combine(sums.data_ptr(), coef.data_ptr(), n,
                             out_disp.data_ptr(), out_err.data_ptr(), stream),
          'usl_loss_combine')
    return out_disp, out_err, sums


_COEF_CACHE = {}
_REDUCE_STREAMS = {}


def _reduce_stream(device) -> 'torch.cuda.Stream':
    """Per-device side stream on which the term sums are reduced, exchanged
    between ranks (sharded batch) and combined beside the transposed warps
    (usl_loss_grad_sharded)."""
    key = (device.type, device.index)
    st = _REDUCE_STREAMS.get(key)
    if st is None:
        # (high priority: its few CTAs go ahead of the pending CTAs of the
        #  transposed-warp kernels)
        st = _REDUCE_STREAMS[key] = torch.cuda.Stream(device, priority=-1)
    return st


def coef_tensor(coefs, device) -> Tensor:
    """Device copy of the per-scale coefficient table (cached: one H2D copy per
    distinct configuration/shape, not per step)."""
    key = (coefs, str(device))
    t = _COEF_CACHE.get(key)
    if t is None:
        t = torch.tensor(coefs, dtype=torch.float32, device=device)
        if len(_COEF_CACHE) > 256:
            _COEF_CACHE.clear()
        _COEF_CACHE[key] = t
    return t


def loss_backward(cfgs: Sequence[UslLossConfig],
                  scales: Sequence[UslLossScale], g_disp: Optional[Tensor],
                  g_err: Optional[Tensor], device, stages: int = 3) -> None:
    L = lib()
    stream = torch.cuda.current_stream(device).cuda_stream
    check(L.usl_loss_bwd(_array(UslLossConfig, cfgs),
                         _array(UslLossScale, scales), len(scales),
                         _ptr(g_disp), _ptr(g_err), stages, stream),
          'usl_loss_bwd')


def loss_regrad(arrays, n_scales: int, g_disp: Tensor, g_err: Tensor,
                device, skip_if_unit: bool = True) -> None:
    """Backward half of the one-pass scheme (`arrays`: the launch description
    of the forward): the gradients written by the
    forward are exact for unit upstream gradients; this launch recomputes them
    for any other upstream pair and returns at once (on the device -- no host
    synchronisation) when both are 1."""
    stream = torch.cuda.current_stream(device).cuda_stream
    cfg_arr, sc_arr = arrays
    check(lib().usl_loss_grad(cfg_arr, sc_arr, n_scales,
                              g_disp.data_ptr(), g_err.data_ptr(), None,
                              GRAD_SKIP_IF_UNIT if skip_if_unit else 0,
                              stream), 'usl_loss_grad')


def grad_rescale(g: Tensor, buffers: Sequence[Tensor], device) -> None:
    """buffers *= g in place unless g == 1 (include/usl.h: usl_grad_rescale)."""
    n = len(buffers)
    ptrs = (C.c_void_p * n)(*[b.data_ptr() for b in buffers])
    counts = (C.c_longlong * n)(*[b.numel() for b in buffers])
    check(lib().usl_grad_rescale(
        g.data_ptr(), ptrs, counts, n,
        torch.cuda.current_stream(device).cuda_stream), 'usl_grad_rescale')


def pyramid(x: Tensor, scales: int) -> List[Tensor]:
    """train/utils.py:27-50.  Level 0 is `x` itself (the reference's level 0
    is a bit-identical copy); levels >= 1 come from one launch."""
    require_cuda_f32(x, 'x')
    x = planes(x)
    b, c, h, w = x.shape
    if scales < 1 or scales > _lib.USL_MAX_SCALES:
        raise ValueError(f'scales must be in [1, {_lib.USL_MAX_SCALES}]')
    out = [x]
    for i in range(1, scales):
        if (h >> i) < 1 or (w >> i) < 1:
            raise ValueError('image too small for the requested scales')
        out.append(torch.empty(b, c, h >> i, w >> i, dtype=x.dtype,
                               device=x.device))
    ptrs = (C.c_void_p * scales)(*[t.data_ptr() for t in out])
    check(lib().usl_pyramid(x.data_ptr(), b, c, h, w, x.stride(0), x.stride(1),
                            scales, ptrs, _stream(x)), 'usl_pyramid')
    return out


def warp_forward(disp: Tensor, sign: float, image: Tensor,
                 out: Tensor) -> None:
    b, c, h, w = image.shape
    check(lib().usl_warp_fwd(disp.data_ptr(), disp.stride(0), sign,
                             image.data_ptr(), image.stride(0),
                             image.stride(1), b, c, h, w, out.data_ptr(),
                             out.stride(0), out.stride(1), _stream(image)),
          'usl_warp_fwd')


def warp_backward_disp(disp: Tensor, sign: float, image: Tensor,
                       grad_out: Tensor, grad_disp: Tensor) -> None:
    b, c, h, w = image.shape
    check(lib().usl_warp_bwd_disp(
        disp.data_ptr(), disp.stride(0), sign, image.data_ptr(),
        image.stride(0), image.stride(1), grad_out.data_ptr(),
        grad_out.stride(0), grad_out.stride(1), b, c, h, w,
        grad_disp.data_ptr(), grad_disp.stride(0), _stream(image)),
        'usl_warp_bwd_disp')


def warp_backward_image(disp: Tensor, sign: float, grad_out: Tensor,
                        grad_image: Tensor) -> None:
    b, c, h, w = grad_out.shape
    nbytes = lib().usl_warp_bwd_image_workspace_bytes(b, c, h, w)
    ws = torch.empty(nbytes // 4, dtype=torch.float32, device=grad_out.device)
    check(lib().usl_warp_bwd_image(
        disp.data_ptr(), disp.stride(0), sign, grad_out.data_ptr(),
        grad_out.stride(0), grad_out.stride(1), b, c, h, w, ws.data_ptr(),
        grad_image.data_ptr(), grad_image.stride(0), grad_image.stride(1),
        _stream(grad_out)), 'usl_warp_bwd_image')


class Reconstruct(torch.autograd.Function):
    """train/utils.py:65-109 as one op: out = warp(image; sign * disp).

    Differentiable w.r.t. the disparity (a deterministic gather) and, like the
    reference's grid_sample, w.r.t. the sampled image (its transpose, also
    deterministic; not a hot path -- in the training step the sampled image is
    data, and the consistency terms, the only place the reference
    differentiates through a sampled map, are handled inside the fused
    loss)."""

    @staticmethod
    def forward(ctx, disp: Tensor, image: Tensor, sign: float) -> Tensor:
        require_cuda_f32(disp, 'disparity')
        require_cuda_f32(image, 'opposite_image')
        if disp.dim() != 4 or disp.size(1) != 1 or \
                disp.shape[-2:] != image.shape[-2:] or \
                disp.size(0) != image.size(0):
            raise ValueError('disparity must be (B,1,h,w) matching the image')
        disp, image = planes(disp), planes(image)
        out = torch.empty(image.shape, dtype=image.dtype, device=image.device)
        warp_forward(disp, sign, image, out)
        ctx.sign = sign
        ctx.save_for_backward(disp, image)
        return out

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        disp, image = ctx.saved_tensors
        grad_out = planes(grad_out)
        grad_disp = grad_image = None
        if ctx.needs_input_grad[0]:
            grad_disp = torch.empty_like(disp, memory_format=torch.contiguous_format)
            warp_backward_disp(disp, ctx.sign, image, grad_out, grad_disp)
        if ctx.needs_input_grad[1]:
            grad_image = torch.empty_like(image, memory_format=torch.contiguous_format)
            warp_backward_image(disp, ctx.sign, grad_out, grad_image)
        return grad_disp, grad_image, None


class ReconstructPair(torch.autograd.Function):
    """One level of train/utils.py:112-135: (B,6,h,w) reconstruction of both
    views from prediction[:, :2] and the stereo pair, written straight into one
    buffer (no torch.cat)."""

    @staticmethod
    def forward(ctx, pred: Tensor, images: Tensor) -> Tensor:
        require_cuda_f32(pred, 'disparity')
        require_cuda_f32(images, 'pyramid level')
        if images.size(1) != 6 or pred.size(1) < 2:
            raise ValueError('expected (B,6,h,w) images and >=2 disparity '
                             'channels')
        pred, images = planes(pred), planes(images)
        out = torch.empty(images.shape, dtype=images.dtype,
                          device=images.device)
        warp_forward(pred[:, 0:1], -1.0, images[:, 3:6], out[:, 0:3])
        warp_forward(pred[:, 1:2], 1.0, images[:, 0:3], out[:, 3:6])
        ctx.save_for_backward(pred, images)
        return out

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        pred, images = ctx.saved_tensors
        if ctx.needs_input_grad[1]:
            raise NotImplementedError('images are data on this path')
        grad_pred = None
        if ctx.needs_input_grad[0]:
            grad_out = planes(grad_out)
            grad_pred = torch.zeros_like(pred,
                                         memory_format=torch.contiguous_format)
            warp_backward_disp(pred[:, 0:1], -1.0, images[:, 3:6],
                               grad_out[:, 0:3], grad_pred[:, 0:1])
            warp_backward_disp(pred[:, 1:2], 1.0, images[:, 0:3],
                               grad_out[:, 3:6], grad_pred[:, 1:2])
        return grad_pred, None


class Pool3(torch.autograd.Function):
    """3x3 valid mean (train/loss.py:386-387) and its transpose."""

    @staticmethod
    def forward(ctx, x: Tensor) -> Tensor:
        require_cuda_f32(x, 'x')
        x = planes(x)
        b, c, h, w = x.shape
        out = torch.empty(b, c, h - 2, w - 2, dtype=x.dtype, device=x.device)
        check(lib().usl_pool3_fwd(x.data_ptr(), x.stride(0), x.stride(1), b, c,
                                  h, w, out.data_ptr(), _stream(x)),
              'usl_pool3_fwd')
        ctx.shape = (b, c, h, w)
        return out

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        b, c, h, w = ctx.shape
        grad_out = grad_out.contiguous()
        gx = torch.empty(b, c, h, w, dtype=grad_out.dtype,
                         device=grad_out.device)
        check(lib().usl_pool3_bwd(grad_out.data_ptr(), b, c, h, w,
                                  gx.data_ptr(), _stream(gx)), 'usl_pool3_bwd')
        return gx


# --------------------------------------------------------------------------
# the fused multi-scale loss as one autograd node
# --------------------------------------------------------------------------
@dataclass
class ScaleSpec:
    """What one scale of a fused call consists of: indices into the flat
    tensor argument list of `FusedLoss.apply` (-1 = absent) and, for the
    two-channel disparity / uncertainty maps, the first channel inside that
    tensor (so a (B,4,h,w) prediction is passed once, without slicing)."""
    terms: int
    coefs: Tuple[float, ...]
    images: int = -1
    disp: int = -1
    disp_ch: int = 0
    unc: int = -1
    unc_ch: int = 0
    recon: int = -1         # given reconstruction (B,6,h,w)
    err: int = -1           # given error map (B,2,h,w)
    want_err: bool = False  # also return the (B,2,h,w) error map
    want_recon: bool = False  # also return the (B,6,h,w) reconstruction, as a
                              # differentiable output (adversarial step)
    flags: int = 0          # USL_SCALE_* (e.g. keep to the general kernels)


def _pair(t: Tensor, ch: int) -> ChannelPair:
    return ChannelPair(t, ch)


class ReconOutputUnavailable(_lib.UslError):
    """The fused kernels cannot write the reconstructions for this call."""


class FusedLoss(torch.autograd.Function):
    """forward(settings, specs, reduce, *tensors) ->
           (disp_loss, error_loss, sums[n,6], error maps that were asked for...,
            reconstructions that were asked for...)

    Gradients are produced for the disparity / uncertainty tensors and for
    given reconstructions; images and given error maps are data
    (loss.py:418 detaches the error)."""

    @staticmethod
    def _grad_buffers(specs, tensors):
        """Fresh gradient tensors for every tensor the kernels write into."""
        covered = [set() for _ in tensors]
        for sp in specs:
            if sp.disp >= 0:
                covered[sp.disp].update((sp.disp_ch, sp.disp_ch + 1))
            if sp.unc >= 0:
                covered[sp.unc].update((sp.unc_ch, sp.unc_ch + 1))
            if sp.recon >= 0:
                covered[sp.recon].update(range(6))
        grads: List[Optional[Tensor]] = [None] * len(tensors)
        for i, t in enumerate(tensors):
            if not covered[i]:
                continue
            alloc = torch.empty_like if len(covered[i]) == t.size(1) \
                else torch.zeros_like
            grads[i] = alloc(t, memory_format=torch.contiguous_format)
        return grads

    @staticmethod
    def _build(settings, specs, tensors, device, grads=None, errs=None,
               keep=None, recons=None):
        """`keep`: list that receives the workspaces the launch description
        points to (they must outlive it)."""
        cfgs, scales = [], []
        for sp in specs:
            b, h, w = _spec_shape(sp, tensors)
            ws = None
            if grads is not None and keep is not None and \
                    sp.terms & (TERM_CONS_D | TERM_CONS_U):
                # {d, u, signed coefficients of the two consistency terms} per pixel
                # for the transposed warp (include/usl.h: scatter_ws)
                ws = torch.empty(b * 2 * h * w * 4, dtype=torch.float32,
                                 device=device)
                keep.append(ws)
            err_out = None
            if errs is not None and sp.want_err:
                err_out = torch.empty(b, 2, h, w, dtype=torch.float32,
                                      device=device)
                errs.append(err_out)
            recon_out = None
            if recons is not None and sp.want_recon:
                recon_out = torch.empty(b, 6, h, w, dtype=torch.float32,
                                        device=device)
                recons.append(recon_out)
            g = grads if grads is not None else [None] * len(tensors)
            cfgs.append(make_config(sp.terms, settings, sp.coefs))
            scales.append(make_scale(
                tensors[sp.images] if sp.images >= 0 else None,
                _pair(tensors[sp.disp], sp.disp_ch) if sp.disp >= 0 else None,
                _pair(tensors[sp.unc], sp.unc_ch) if sp.unc >= 0 else None,
                shape=(b, h, w),
                recon_in=tensors[sp.recon] if sp.recon >= 0 else None,
                err_in=tensors[sp.err] if sp.err >= 0 else None,
                err_out=err_out, recon_out=recon_out,
                grad_disp=_pair(g[sp.disp], sp.disp_ch)
                if sp.disp >= 0 and g[sp.disp] is not None else None,
                grad_unc=_pair(g[sp.unc], sp.unc_ch)
                if sp.unc >= 0 and g[sp.unc] is not None else None,
                grad_recon_out=g[sp.recon] if sp.recon >= 0 else None,
                scatter_ws=ws, flags=sp.flags))
        return cfgs, scales

    @staticmethod
    def forward(ctx, settings: LossSettings, specs: Sequence[ScaleSpec],
                reduce: Optional[Reduce], *tensors: Tensor):
        for i, t in enumerate(tensors):
            require_cuda_f32(t, f'tensor {i}')
        tensors = tuple(planes(t) for t in tensors)
        device = tensors[0].device
        ctx.settings, ctx.specs = settings, specs
        ctx.onepass = None
        errs: List[Tensor] = []
        recons: List[Tensor] = []
        ctx.n_recon = 0
        if any(ctx.needs_input_grad[3:]):
            # one pass: the sums and (for unit upstream gradients) the
            # gradients together -- see usl_loss_grad in include/usl.h
            grads = FusedLoss._grad_buffers(specs, tensors)
            keep: List[Tensor] = []
            cfgs, scales = FusedLoss._build(settings, specs, tensors, device,
                                            grads, errs, keep, recons)
            n = len(scales)
            arrays = (_array(UslLossConfig, cfgs), _array(UslLossScale, scales))
            starts = plan_rows(arrays[0], arrays[1], n, MODE_GRAD)
            if starts is not None:
                out_disp, out_err, sums = loss_forward(
                    cfgs, scales, device, reduce, with_grad=True,
                    arrays=arrays, starts=starts)
                ctx.onepass = grads
                ctx.workspaces = keep
                ctx.backward_calls = 0
                # the same launch description serves the backward: the tensors
                # it points to are kept alive by save_for_backward / `grads`;
                # the error maps are outputs the caller may drop, so the
                # backward must not write them again
                for i in range(n):
                    arrays[1][i].err_out = None
                    arrays[1][i].recon_out = None
                ctx.arrays = arrays
                ctx.save_for_backward(*tensors)
                ctx.mark_non_differentiable(sums, *errs)
                ctx.set_materialize_grads(False)
                ctx.n_recon = len(recons)
                return (out_disp, out_err, sums) + tuple(errs) + tuple(recons)
            errs = []
            if any(sp.want_recon for sp in specs):
                raise ReconOutputUnavailable(
                    'reconstruction outputs need the one-pass launch')
        # forward only (no gradient asked for): sums, and the optional maps
        recons = []
        cfgs, scales = FusedLoss._build(settings, specs, tensors, device,
                                        None, errs, recons=recons)
        out_disp, out_err, sums = loss_forward(cfgs, scales, device, reduce)
        ctx.save_for_backward(*tensors)
        ctx.mark_non_differentiable(sums, *errs, *recons)
        ctx.set_materialize_grads(False)
        return (out_disp, out_err, sums) + tuple(errs) + tuple(recons)

    @staticmethod
    def _recon_backward(ctx, specs, tensors, grads, g_recons, device):
        """A gradient arriving at the reconstruction outputs (from a
        discriminator) -> added to the disparity gradients, one launch."""
        levels, gptr, optr, obs, ocs, keep = [], [], [], [], [], []
        k = 0
        for sp in specs:
            if not sp.want_recon:
                continue
            g = g_recons[k]
            k += 1
            if g is None:
                continue
            g = g.contiguous()
            im, pr, gp = tensors[sp.images], tensors[sp.disp], grads[sp.disp]
            lv = _lib.UslDiscLevel()
            lv.B, lv.h, lv.w = im.size(0), im.size(2), im.size(3)
            lv.images, lv.img_bs, lv.img_cs = im.data_ptr(), im.stride(0), im.stride(1)
            lv.pred = pr.data_ptr() + 4 * sp.disp_ch * pr.stride(1)
            lv.pred_bs, lv.pred_cs = pr.stride(0), pr.stride(1)
            levels.append(lv)
            gptr.append(g.data_ptr())
            optr.append(gp.data_ptr() + 4 * sp.disp_ch * gp.stride(1))
            obs.append(gp.stride(0)); ocs.append(gp.stride(1))
            keep.append(g)
        if not levels:
            return
        n = len(levels)
        check(lib().usl_recon_bwd(
            _array(_lib.UslDiscLevel, levels), (C.c_void_p * n)(*gptr),
            (C.c_void_p * n)(*optr), (C.c_longlong * n)(*obs),
            (C.c_longlong * n)(*ocs), n, 1,
            torch.cuda.current_stream(device).cuda_stream), 'usl_recon_bwd')

    @staticmethod
    def backward(ctx, g_disp, g_err, *unused):
        tensors = ctx.saved_tensors
        specs, settings = ctx.specs, ctx.settings
        device = tensors[0].device
        if g_disp is not None:
            g_disp = g_disp.contiguous()
        if g_err is not None:
            g_err = g_err.contiguous()
        needs = ctx.needs_input_grad[3:]
        if ctx.onepass is not None:
            zero = None
            if g_disp is None or g_err is None:
                zero = torch.zeros((), dtype=torch.float32, device=device)
            gd = zero if g_disp is None else g_disp
            ge = zero if g_err is None else g_err
            ctx.backward_calls += 1
            if ctx.backward_calls == 1:
                # the buffers hold the gradients for unit upstream gradients
                # (written by the forward)
                grads = ctx.onepass
                live = [g for g in grads if g is not None]
                if gd.data_ptr() == ge.data_ptr() and \
                        len(live) <= _lib.USL_MAX_RESCALE:
                    # one upstream gradient for both outputs -- what
                    # (disp_loss + error_loss).backward() hands over: scale in
                    # place unless it is 1 (decided on the device), one launch
                    grad_rescale(gd, live, device)
                else:
                    # redone in place if (gd, ge) differ from (1, 1)
                    loss_regrad(ctx.arrays, len(specs), gd, ge, device, True)
            else:
                # backward(retain_graph=True) again: the first call's buffers
                # were handed to autograd (and may hold non-unit gradients
                # now), so this call computes into fresh ones, unconditionally
                grads = FusedLoss._grad_buffers(specs, tensors)
                keep = []
                cfgs, scales = FusedLoss._build(settings, specs, tensors,
                                                device, grads, keep=keep)
                arrays = (_array(UslLossConfig, cfgs),
                          _array(UslLossScale, scales))
                loss_regrad(arrays, len(specs), gd, ge, device, False)
            if ctx.n_recon:
                FusedLoss._recon_backward(ctx, specs, tensors, grads,
                                          unused[len(unused) - ctx.n_recon:],
                                          device)
            return (None, None, None) + tuple(
                g if need else None for g, need in zip(grads, needs))
        grads = FusedLoss._grad_buffers(specs, tensors)
        keep = []
        cfgs, scales = FusedLoss._build(settings, specs, tensors, device,
                                        grads, keep=keep)
        loss_backward(cfgs, scales, g_disp, g_err, device)
        return (None, None, None) + tuple(
            g if need else None for g, need in zip(grads, needs))


def _spec_shape(sp: ScaleSpec, tensors) -> Tuple[int, int, int]:
    for idx in (sp.images, sp.disp, sp.unc, sp.err):
        if idx >= 0:
            b, _, h, w = tensors[idx].shape
            return b, h, w
    raise ValueError('empty scale spec')


# --------------------------------------------------------------------------
# either side of the loss path
# --------------------------------------------------------------------------
def _ptr_array(tensors: Sequence[Tensor]):
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def _count_array(tensors: Sequence[Tensor]):
    return (C.c_longlong * len(tensors))(*[t.numel() for t in tensors])


class DisparityHead(torch.autograd.Function):
    """The decoder's output activation over a whole pyramid (reference
    model/layers/decoder.py:239-246): pred_i = scale * sigmoid(logits_i), every
    level in one launch, forward and backward."""

    @staticmethod
    def forward(ctx, scale: float, *logits: Tensor):
        for i, t in enumerate(logits):
            require_cuda_f32(t, f'logits {i}')
        if not 0 < len(logits) <= _lib.USL_MAX_SCALES:
            raise ValueError(f'1 to {_lib.USL_MAX_SCALES} levels')
        if not scale > 0:
            raise ValueError('scale must be positive')
        logits = [t.contiguous() for t in logits]
        preds = [torch.empty_like(t) for t in logits]
        check(lib().usl_head_fwd(_ptr_array(logits), _ptr_array(preds),
                                 _count_array(logits), len(logits),
                                 float(scale), _stream(logits[0])),
              'usl_head_fwd')
        ctx.scale = float(scale)
        ctx.save_for_backward(*preds)
        return tuple(preds)

    @staticmethod
    def backward(ctx, *grads: Tensor):
        preds = ctx.saved_tensors
        gin = [torch.zeros_like(p) if g is None else g.contiguous()
               for g, p in zip(grads, preds)]
        gout = [torch.empty_like(p) for p in preds]
        check(lib().usl_head_bwd(_ptr_array(gin), _ptr_array(preds),
                                 _ptr_array(gout), _count_array(preds),
                                 len(preds), ctx.scale, _stream(preds[0])),
              'usl_head_bwd')
        return (None,) + tuple(gout)


def disc_input(images: Sequence[Tensor], preds: Optional[Sequence[Tensor]],
               recons: Optional[Sequence[Tensor]]) -> List[Tensor]:
    """[image level ; its reconstruction] along the batch axis for every level,
    one launch (include/usl.h: usl_disc_input).  `preds` given: the
    reconstruction half is warped on the fly; else copied from `recons`."""
    n = len(images)
    lv = (_lib.UslDiscLevel * n)()
    outs, keep = [], []
    for i in range(n):
        im = planes(images[i])
        require_cuda_f32(im, f'image level {i}')
        b, c, h, w = im.shape
        if c != 6:
            raise ValueError('image levels must have 6 channels')
        out = torch.empty(2 * b, 6, h, w, dtype=im.dtype, device=im.device)
        lv[i].B, lv[i].h, lv[i].w = b, h, w
        lv[i].images, lv[i].img_bs, lv[i].img_cs = im.data_ptr(), im.stride(0), im.stride(1)
        if preds is not None:
            p = planes(preds[i].detach())
            require_cuda_f32(p, f'prediction level {i}')
            lv[i].pred, lv[i].pred_bs, lv[i].pred_cs = p.data_ptr(), p.stride(0), p.stride(1)
            keep.append(p)
        else:
            r = planes(recons[i].detach())
            require_cuda_f32(r, f'reconstruction level {i}')
            lv[i].recon, lv[i].rec_bs, lv[i].rec_cs = r.data_ptr(), r.stride(0), r.stride(1)
            keep.append(r)
        lv[i].out = out.data_ptr()
        outs.append(out)
        keep.append(im)
    check(lib().usl_disc_input(lv, n, _stream(outs[0])), 'usl_disc_input')
    return outs
