"""Pyramid and warp helpers -- drop-in for the hot-path part of the
reference's `train/utils.py` (lines 17-140).  Same names, same argument
meaning; the work is done by libusl.so on the GPU.
"""
import ctypes as C
from typing import List, Optional, Sequence, Union

import torch
from torch import Tensor
from torch.nn import Module

from .. import functional as K
from .._lib import check, lib

Device = Union[torch.device, str]
ImagePyramid = List[Tensor]


def l1_loss(x: Tensor, y: Tensor) -> Tensor:
    """Mean absolute difference (reference utils.py:22-24).  Used on
    discriminator feature maps of arbitrary shape, so it stays a torch op."""
    return torch.mean(torch.abs(x - y))


def scale_pyramid(x: Tensor, scales: int) -> ImagePyramid:
    """Bilinear (align_corners=True) image pyramid, every level resampled
    from the full-resolution input (reference utils.py:27-50).

    Level 0 is `x` itself, not a copy: the reference's level 0 is a
    bit-identical resample.  Nothing on the path writes into a level; a caller
    who mutates level 0 in place mutates `x` (clone it first if that matters).
    One kernel launch produces all other levels."""
    return K.pyramid(x, scales)


def detach_pyramid(pyramid: Sequence[Tensor]) -> ImagePyramid:
    """Copy of a pyramid cut from the autograd graph (utils.py:53-62)."""
    return [level.detach().clone() for level in pyramid]


def reconstruct(disparity: Tensor, opposite_image: Tensor) -> Tensor:
    """Warp `opposite_image` horizontally by `disparity` (in units of the
    image width), reference utils.py:65-97."""
    return K.Reconstruct.apply(disparity, opposite_image, 1.0)


def reconstruct_left_image(left_disparity: Tensor,
                           right_image: Tensor) -> Tensor:
    """Left view from the right image (utils.py:100-103): shift = -disparity."""
    return K.Reconstruct.apply(left_disparity, right_image, -1.0)


def reconstruct_right_image(right_disparity: Tensor,
                            left_image: Tensor) -> Tensor:
    """Right view from the left image (utils.py:106-109): shift = +disparity."""
    return K.Reconstruct.apply(right_disparity, left_image, 1.0)


class ReconPyramid(Sequence):
    """What `reconstruct_pyramid` returns: a pyramid of reconstructions that
    is only materialised when somebody looks at it.

    The training step (reference train.py:122-124) builds the reconstruction
    pyramid and hands it straight to the loss.  The fused loss kernels warp
    in-kernel, so when `TukraUncertaintyLoss` receives this object untouched
    it never writes the reconstructions to HBM.  Any other consumer (the
    discriminator, `detach_pyramid`, indexing, iteration) materialises the
    levels -- as differentiable tensors -- on first access, and the loss then
    uses them as given.
    """

    def __init__(self, disparities: Sequence[Tensor],
                 pyramid: Sequence[Tensor]) -> None:
        self.disparities = list(disparities)
        self.pyramid = list(pyramid)
        if len(self.disparities) != len(self.pyramid):
            # zip() semantics of the reference: the shorter one wins
            n = min(len(self.disparities), len(self.pyramid))
            self.disparities = self.disparities[:n]
            self.pyramid = self.pyramid[:n]
        self._levels: Optional[List[Tensor]] = None

    @property
    def materialised(self) -> bool:
        return self._levels is not None

    def tensors(self) -> List[Tensor]:
        if self._levels is None:
            self._levels = [K.ReconstructPair.apply(d, im)
                            for d, im in zip(self.disparities, self.pyramid)]
        return self._levels

    def adopt(self, levels: Sequence[Tensor]) -> None:
        """The levels, computed elsewhere (the fused loss kernels write them
        out in the adversarial step)."""
        if len(levels) != len(self.disparities):
            raise ValueError('one reconstruction per level')
        self._levels = list(levels)

    def built_from(self, disparities: Sequence[Tensor],
                   pyramid: Sequence[Tensor]) -> bool:
        return len(disparities) == len(self.disparities) and \
            len(pyramid) == len(self.pyramid) and \
            all(a is b for a, b in zip(disparities, self.disparities)) and \
            all(a is b for a, b in zip(pyramid, self.pyramid))

    def __len__(self) -> int:
        return len(self.disparities)

    def __getitem__(self, i):
        return self.tensors()[i]

    def __iter__(self):
        return iter(self.tensors())


def reconstruct_pyramid(disparities: Sequence[Tensor],
                        pyramid: Sequence[Tensor]) -> ReconPyramid:
    """Reconstruction of both views at every scale from the first two
    prediction channels (reference utils.py:112-135), lazily."""
    return ReconPyramid(disparities, pyramid)


def concatenate_pyramids(a: Sequence[Tensor],
                         b: Sequence[Tensor]) -> ImagePyramid:
    """Level-wise concatenation along the batch axis (utils.py:138-140)."""
    return [torch.cat((x, y), 0) for x, y in zip(a, b)]


def run_discriminator(image_pyramid: Sequence[Tensor],
                      recon_pyramid: Sequence[Tensor], discriminator: Module,
                      disc_loss_function: Module, batch_size: int) -> Tensor:
    """Discriminator predictions on [real ; reconstructed] and its loss
    (reference utils.py:248-273).  The reference detaches + clones the
    reconstruction pyramid and concatenates it with the image pyramid (two
    copies per level); here one kernel writes the concatenated levels -- and
    when `recon_pyramid` is still lazy (nobody looked at it) the
    reconstruction half is warped straight into them."""
    images = list(image_pyramid)
    if isinstance(recon_pyramid, ReconPyramid) and not recon_pyramid.materialised:
        n = min(len(images), len(recon_pyramid))
        pyramid = K.disc_input(images[:n], recon_pyramid.disparities[:n], None)
    else:
        recons = list(recon_pyramid)
        n = min(len(images), len(recons))
        pyramid = K.disc_input(images[:n], None, recons[:n])
    predictions = discriminator(pyramid)
    labels = torch.zeros_like(predictions)
    labels[:batch_size] = 1
    return disc_loss_function(predictions, labels) / 2


def disparity_head(logits: Sequence[Tensor], scale: float) -> List[Tensor]:
    """The decoder's output activation for a pyramid of logits (reference
    model/layers/decoder.py:239-246): `scale * sigmoid(x)` per level, one
    launch for all levels, differentiable."""
    return list(K.DisparityHead.apply(scale, *logits))


_INFERNO = None


def colour_table(colour_map: str = 'inferno') -> Tensor:
    """The (N,3) fp64 RGB table of a matplotlib colour map (utils.py:196
    fetches it with plt.get_cmap)."""
    try:
        import matplotlib.pyplot as plt
    except ImportError as e:
        raise ImportError('to_heatmap(colour_map=<name>) needs matplotlib for '
                          'the colour table; pass the (N,3) table itself '
                          'instead') from e
    cmap = plt.get_cmap(colour_map)
    import numpy as np
    return torch.from_numpy(cmap(np.arange(cmap.N))[:, :3].astype('float64'))


def to_heatmap(x: Tensor, device: Device = 'cpu', inverse: bool = False,
               colour_map: Union[str, Tensor] = 'inferno') -> Tensor:
    """Single-channel image -> RGB heat map (reference utils.py:177-196), on
    the GPU: (1,H,W) fp32 -> (3,H,W) fp64 (matplotlib's tables are fp64).
    `colour_map`: a matplotlib name, or the (N,3) table itself."""
    K.require_cuda_f32(x, 'x')
    table = colour_table(colour_map) if isinstance(colour_map, str) else colour_map
    table = table.to(device=x.device, dtype=torch.float64).contiguous()
    img = x.squeeze(0).contiguous()
    out = torch.empty((3,) + tuple(img.shape), dtype=torch.float64, device=x.device)
    check(lib().usl_heatmap(img.data_ptr(), img.numel(), int(bool(inverse)),
                            table.data_ptr(), table.size(0), out.data_ptr(),
                            torch.cuda.current_stream(x.device).cuda_stream),
          'usl_heatmap')
    return out.to(device)


def combine_disparity(left: Tensor, right: Tensor, device: Device = 'cpu',
                      alpha: float = 20, beta: float = 0.05) -> Tensor:
    """Blend of the two views' disparities that hides each one's blind spot
    (reference utils.py:199-245, after Monodepth2), on the GPU: (C,H,W) fp32
    pair -> (C,H,W) fp64 like the numpy original."""
    K.require_cuda_f32(left, 'left')
    K.require_cuda_f32(right, 'right')
    if left.shape != right.shape or left.dim() != 3:
        raise ValueError('left and right must both be (C,H,W)')
    a, b = left.contiguous(), right.contiguous()
    c, h, w = a.shape
    out = torch.empty(c, h, w, dtype=torch.float64, device=a.device)
    check(lib().usl_combine_disparity(
        a.data_ptr(), b.data_ptr(), c, h, w, float(alpha), float(beta),
        out.data_ptr(), torch.cuda.current_stream(a.device).cuda_stream),
        'usl_combine_disparity')
    return out.to(device)
