"""Loss modules -- drop-in for the reference's `train/loss.py`.

Same class names, constructor kwargs (they are built from `config.yml`,
reference main.py:108), call signatures, return types and error behaviour.
The per-pixel work of every class runs in the fused CUDA kernels of
libusl.so; each class here only decides which terms of the fused kernel are
switched on and with which coefficient.  The modules own no parameters or
buffers (like the reference's), so `.to(device)` / `state_dict()` keep working.

Inputs must be CUDA float32 tensors -- there is no CPU fallback.
"""
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
from torch import Tensor
from torch.nn import Module
from torch.nn.parallel import DistributedDataParallel

from .. import functional as K
from .._lib import (TERM_CONS_D, TERM_CONS_U, TERM_REPROJ, TERM_SMOOTH_D,
                    TERM_SMOOTH_U, TERM_UNC)
from ..functional import FusedLoss, LossSettings, ScaleSpec
from . import utils as u
from .utils import ImagePyramid, ReconPyramid

_ZERO6 = (0.0,) * 6


def _coefs(**kw) -> Tuple[float, ...]:
    names = ('reproj', 'cons_d', 'smooth_d', 'unc', 'smooth_u', 'cons_u')
    return tuple(float(kw.get(n, 0.0)) for n in names)


class WeightedSSIMLoss(nn.Module):
    """alpha * DSSIM(3x3) + (1 - alpha) * L1 between a stereo pair and its
    reconstruction (reference loss.py:15-151).

    Args:
        alpha: weight of the SSIM part (L1 gets 1 - alpha). Default 0.85.
        k1, k2: SSIM stabilisers (squared internally). Defaults 0.01, 0.03.
    """

    def __init__(self, alpha: float = 0.85, k1: float = 0.01,
                 k2: float = 0.03) -> None:
        super().__init__()
        self.alpha = alpha
        self.k1 = k1 ** 2
        self.k2 = k2 ** 2
        self._last_error = None      # tensor, or a thunk producing it

    def _settings(self) -> LossSettings:
        return LossSettings(alpha=self.alpha, c1=self.k1, c2=self.k2)

    @property
    def previous_image_error(self) -> Optional[Tensor]:
        """The (B,2,h,w) error map of the most recent evaluation."""
        if callable(self._last_error):
            self._last_error = self._last_error()
        return self._last_error

    def _run(self, images: Tensor, recon: Tensor, want_err: bool):
        b, _, h, w = images.shape
        spec = ScaleSpec(terms=TERM_REPROJ,
                         coefs=_coefs(reproj=1.0 / (b * h * w)),
                         images=0, recon=1, want_err=want_err)
        return FusedLoss.apply(self._settings(), [spec], None, images, recon)

    def image_error(self, images: Tensor, recon: Tensor) -> Tensor:
        """Per-pixel, per-view error map (B,2,h,w).

        The map itself carries no autograd history (the reference's callers
        use it detached: evaluate.py:151 under no_grad, loss.py:418);
        gradients reach `recon` through `forward`."""
        with torch.no_grad():
            return self._run(images, recon, True)[3]

    def forward(self, images: Tensor, recon: Tensor) -> Tensor:
        """mean over pixels of (left error + right error), loss.py:133-151."""
        out = self._run(images, recon, True)
        self._last_error = out[3]
        return out[0]


class ConsistencyLoss(nn.Module):
    """Left-right consistency (reference loss.py:154-188): each view of
    `disp` is compared with the opposite view of `images` (default: `disp`
    itself) warped by it."""

    def __init__(self) -> None:
        super().__init__()

    def forward(self, disp: Tensor, images: Optional[Tensor] = None) -> Tensor:
        b, _, h, w = disp.shape
        n = b * h * w
        if images is None:
            spec = ScaleSpec(terms=TERM_CONS_D, coefs=_coefs(cons_d=1.0 / n),
                             disp=0)
            return FusedLoss.apply(LossSettings(), [spec], None, disp)[0]
        # shift map in the "uncertainty" slot, sampled map in the "disparity"
        # slot of the kernel (loss.py:430-431 is this very call)
        spec = ScaleSpec(terms=TERM_CONS_U, coefs=_coefs(cons_u=1.0 / n),
                         unc=0, disp=1)
        return FusedLoss.apply(LossSettings(), [spec], None, disp, images)[1]


class SmoothnessLoss(nn.Module):
    """Edge-aware first-order smoothness (reference loss.py:191-264)."""

    def __init__(self) -> None:
        super().__init__()

    def forward(self, disp: Tensor, images: Tensor) -> Tensor:
        b, _, h, w = disp.shape
        spec = ScaleSpec(terms=TERM_SMOOTH_D,
                         coefs=_coefs(smooth_d=1.0 / (b * h * w)),
                         images=1, disp=0)
        return FusedLoss.apply(LossSettings(), [spec], None, disp, images)[0]


class PerceptualLoss(nn.Module):
    """L1 distance between discriminator feature maps of the original and the
    reconstructed pyramid (reference loss.py:267-305).  The discriminator is a
    conv net outside the hot path; this is plain PyTorch."""

    def __init__(self) -> None:
        super().__init__()

    def forward(self, image_pyramid: ImagePyramid,
                recon_pyramid: ImagePyramid, disc: Module) -> Tensor:
        net = disc.module if isinstance(disc, DistributedDataParallel) else disc
        real = net.features(list(image_pyramid))
        fake = net.features(list(recon_pyramid))
        total = 0
        for a, b in zip(real, fake):
            total = total + u.l1_loss(a, b)
        return total


class GeneratorLoss(nn.Module):
    """Adversarial term: the discriminator should call the reconstructions
    real (reference loss.py:308-337).  Plain PyTorch."""

    def __init__(self, loss: str = 'mse') -> None:
        super().__init__()
        self.adversarial = nn.MSELoss() if loss == 'mse' else nn.BCELoss()

    def forward(self, recon_pyramid: ImagePyramid,
                discriminator: Module) -> Tensor:
        verdict = discriminator(list(recon_pyramid))
        return self.adversarial(verdict, torch.ones_like(verdict))


class ReprojectionErrorLoss(nn.Module):
    """Loss of the two uncertainty channels against the (detached)
    reprojection error (reference loss.py:340-434).

    Args:
        loss_type: 'l1' | 'bayesian' (Laplacian NLL, sigma^2 predicted) |
            'log_bayesian' (log sigma^2 predicted). Default 'l1'.
        smoothness_weight: weight of SmoothnessLoss(uncertainty). Default 1.0.
        consistency_weight: weight of ConsistencyLoss(uncertainty, disparity).
            Default 1.0.
        pooling: 3x3 average-pool prediction, image and error first.
    """

    def __init__(self, loss_type: str = 'l1', smoothness_weight: float = 1.0,
                 consistency_weight: float = 1.0,
                 pooling: bool = False) -> None:
        super().__init__()
        if loss_type not in ('l1', 'bayesian', 'log_bayesian'):
            raise ValueError('Loss must be either "l1", "bayesian" '
                             'or "log_bayesian".')
        self.loss_type = loss_type
        self.smoothness_weight = smoothness_weight
        self.consistency_weight = consistency_weight
        self.pooling = bool(pooling)

    def _settings(self) -> LossSettings:
        return LossSettings(loss_type=self.loss_type,
                            err_smoothness_weight=self.smoothness_weight,
                            err_consistency_weight=self.consistency_weight)

    def error_terms(self) -> int:
        t = TERM_UNC
        if self.smoothness_weight > 0:
            t |= TERM_SMOOTH_U
        if self.consistency_weight > 0:
            t |= TERM_CONS_U
        return t

    def forward(self, predicted: Tensor, image: Tensor, error: Tensor) -> Tensor:
        error = error.detach()                      # loss.py:418
        if self.pooling:                            # loss.py:420-422
            predicted = K.Pool3.apply(predicted)
            image = K.Pool3.apply(image)
            error = K.Pool3.apply(error)
        b, _, h, w = predicted.shape
        st = self._settings()
        coefs = st.coefs(0, b * h * w)
        coefs[0] = coefs[1] = coefs[2] = 0.0
        spec = ScaleSpec(terms=self.error_terms(), coefs=tuple(coefs),
                         images=1, disp=0, disp_ch=0, unc=0, unc_ch=2, err=2)
        return FusedLoss.apply(st, [spec], None, predicted, image, error)[1]


class TukraUncertaintyLoss(nn.Module):
    """Total multi-scale loss of the uncertainty model (reference
    loss.py:437-568): per scale, reprojection (WSSIM) + consistency +
    smoothness / 2^i for the disparity, and the uncertainty loss; optionally a
    generator and a perceptual term from a discriminator.

    Constructor arguments are those of the reference (the `loss:` block of
    config.yml).  `forward` returns the tuple (disparity loss, error loss).

    All scales, all terms and both outputs come out of ONE fused forward
    launch (and the backward out of two), provided `recon_pyramid` is the
    untouched result of `utils.reconstruct_pyramid(predictions,
    image_pyramid)`; any other reconstruction pyramid is honoured as given.

    `uncertainty_model_b200.distributed.shard_loss` makes the returned losses
    those of the global batch when the batch is sharded over ranks, with the
    gradients either of the local mean (torch DDP averages them) or of the
    global mean.
    """

    def __init__(self, wssim_weight: float = 1.0,
                 consistency_weight: float = 1.0,
                 smoothness_weight: float = 1.0,
                 adversarial_weight: float = 0.85,
                 predictive_error_weight: float = 1.0,
                 perceptual_weight: float = 0.05,
                 wssim_alpha: float = 0.85,
                 perceptual_start: int = 5,
                 adversarial_loss_type: str = 'mse',
                 error_loss_config: Optional[dict] = None) -> None:
        super().__init__()
        self.wssim = WeightedSSIMLoss(wssim_alpha)
        self.consistency = ConsistencyLoss()
        self.smoothness = SmoothnessLoss()
        self.adversarial = GeneratorLoss(adversarial_loss_type)
        self.perceptual = PerceptualLoss()
        self.predictive_error = ReprojectionErrorLoss(
            **(error_loss_config or {}))

        self.perceptual_start = perceptual_start
        self.wssim_weight = wssim_weight
        self.consistency_weight = consistency_weight
        self.smoothness_weight = smoothness_weight
        self.adversarial_weight = adversarial_weight
        self.perceptual_weight = perceptual_weight
        self.predictive_error_weight = predictive_error_weight

        # batch sharded over ranks (uncertainty_model_b200.distributed):
        self.reduce_group = None     # torch.distributed group, or None
        self.world_size = 1          # ranks the reported losses are means over
        self.grad_world_size = 1     # ... the gradients are normalised by
        self.last_term_sums = None   # fp64 [scales, 6] raw sums (debugging)
        self.kernel_flags = 0        # USL_SCALE_* of include/usl.h (tests)

    def _settings(self) -> LossSettings:
        pe = self.predictive_error
        return LossSettings(
            wssim_weight=self.wssim_weight,
            consistency_weight=self.consistency_weight,
            smoothness_weight=self.smoothness_weight,
            predictive_error_weight=self.predictive_error_weight,
            alpha=self.wssim.alpha, c1=self.wssim.k1, c2=self.wssim.k2,
            loss_type=pe.loss_type,
            err_smoothness_weight=pe.smoothness_weight,
            err_consistency_weight=pe.consistency_weight)

    def forward(self, image_pyramid: ImagePyramid, predictions: ImagePyramid,
                recon_pyramid: Sequence[Tensor], epoch: Optional[int] = None,
                discriminator: Optional[Module] = None
                ) -> Tuple[Tensor, Tensor]:
        n_scales = min(len(image_pyramid), len(predictions),
                       len(recon_pyramid))
        images = list(image_pyramid)[:n_scales]
        preds = list(predictions)[:n_scales]

        fused = isinstance(recon_pyramid, ReconPyramid) \
            and not recon_pyramid.materialised \
            and recon_pyramid.built_from(list(predictions),
                                         list(image_pyramid))
        # adversarial step on a still-lazy pyramid: the fused kernels warp
        # in-kernel as always and write the reconstructions out as well (a
        # differentiable output the discriminator terms consume)
        emit_recon = fused and discriminator is not None \
            and not self.predictive_error.pooling \
            and all(t.requires_grad for t in preds)
        if fused and discriminator is not None and not emit_recon:
            fused = False
        recons: Optional[List[Tensor]] = None
        if not fused:
            recons = list(recon_pyramid)[:n_scales]

        st = self._settings()
        pooling = self.predictive_error.pooling
        disp_terms = TERM_REPROJ | TERM_CONS_D | TERM_SMOOTH_D
        err_terms = self.predictive_error.error_terms()

        tensors: List[Tensor] = []
        specs: List[ScaleSpec] = []
        for i in range(n_scales):
            b, _, h, w = preds[i].shape
            coefs = st.coefs(i, b * self.grad_world_size * h * w)
            if pooling:
                coefs[3] = coefs[4] = coefs[5] = 0.0
            sp = ScaleSpec(terms=disp_terms if pooling
                           else disp_terms | err_terms,
                           coefs=tuple(coefs), want_err=pooling,
                           want_recon=emit_recon, flags=self.kernel_flags)
            sp.images = len(tensors); tensors.append(images[i])
            sp.disp = sp.unc = len(tensors); tensors.append(preds[i])
            sp.disp_ch, sp.unc_ch = 0, 2
            if recons is not None:
                sp.recon = len(tensors); tensors.append(recons[i])
            specs.append(sp)

        reduce = None
        if self.reduce_group is not None:
            reduce = K.Reduce(self.reduce_group,
                              self.grad_world_size / self.world_size)
        try:
            out = FusedLoss.apply(st, specs, reduce, *tensors)
        except K.ReconOutputUnavailable:
            # (shapes the one-pass kernels do not take: materialise the
            #  pyramid with the stand-alone warp and honour it as given)
            recon_pyramid.tensors()
            return self.forward(image_pyramid, predictions, recon_pyramid,
                                epoch, discriminator)
        disp_loss, error_loss, sums = out[0], out[1], out[2]
        self.last_term_sums = sums
        if emit_recon:
            recons = list(out[3:3 + n_scales])
            # whoever looks at the pyramid next (run_discriminator) finds it
            recon_pyramid.adopt(recons)

        if pooling:
            # loss.py:420-422: the error terms run on 3x3-pooled copies
            errs = out[3:]
            p_tensors: List[Tensor] = []
            p_specs: List[ScaleSpec] = []
            for i in range(n_scales):
                pp = K.Pool3.apply(preds[i])
                pi = K.Pool3.apply(images[i])
                pe = K.Pool3.apply(errs[i])
                b, _, h, w = pp.shape
                coefs = st.coefs(i, b * self.grad_world_size * h * w)
                coefs[0] = coefs[1] = coefs[2] = 0.0
                sp = ScaleSpec(terms=err_terms, coefs=tuple(coefs))
                sp.images = len(p_tensors); p_tensors.append(pi)
                sp.disp = sp.unc = len(p_tensors); p_tensors.append(pp)
                sp.disp_ch, sp.unc_ch = 0, 2
                sp.err = len(p_tensors); p_tensors.append(pe)
                p_specs.append(sp)
            error_loss = FusedLoss.apply(st, p_specs, reduce,
                                         *p_tensors)[1]
            self.wssim._last_error = errs[-1]
        else:
            self.wssim._last_error = self._error_thunk(
                images[-1], preds[-1], None if recons is None else recons[-1])

        if discriminator is not None:
            # loss.py:552-558 -- conv-net consumers of the reconstructions
            disp_loss = disp_loss + self.adversarial_weight * \
                self.adversarial(recons, discriminator)
            if epoch is not None and epoch >= self.perceptual_start:
                disp_loss = disp_loss + self.perceptual_weight * \
                    self.perceptual(images, recons, discriminator)

        return disp_loss, error_loss

    def _error_thunk(self, images: Tensor, pred: Tensor,
                     recon: Optional[Tensor]):
        """`self.wssim.previous_image_error` (loss.py:548) on demand: the fused
        step never writes the error map to HBM unless somebody asks."""
        images, pred = images.detach(), pred.detach()
        recon = None if recon is None else recon.detach()
        settings = self._settings()

        def compute() -> Tensor:
            with torch.no_grad():
                rc = recon if recon is not None \
                    else K.ReconstructPair.apply(pred, images)
                b, _, h, w = images.shape
                spec = ScaleSpec(terms=TERM_REPROJ,
                                 coefs=_coefs(reproj=1.0 / (b * h * w)),
                                 images=0, recon=1, want_err=True)
                return FusedLoss.apply(settings, [spec], None, images, rc)[3]
        return compute
