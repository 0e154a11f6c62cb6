"""Pyramid and warp helpers -- drop-in for the hot-path part of the
reference's `train/utils.py` (lines 17-140).  Same names, same argument
meaning; the work is done by libusl.so on the GPU.
"""
from typing import List, Optional, Sequence, Union

import torch
from torch import Tensor

from .. import functional as K

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
    bit-identical resample.  One kernel launch produces all other levels."""
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
