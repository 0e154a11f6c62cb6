"""TEST INFRASTRUCTURE ONLY -- CPU oracles for the callers either side of the
loss path: the decoder's disparity head, the
discriminator input glue and the evaluation post-processing.

Only `tests/` may import this module.  All `file:line` citations are relative
to the reference tree.
"""
from typing import List, Sequence

import numpy as np
import torch
from torch import Tensor


def disparity_head(logits: Sequence[Tensor], scale: float) -> List[Tensor]:
    """model/layers/decoder.py:239-246: `scale * self.disp(out)` with
    `self.disp = ConvLayer(..., sigmoid=True)` -- the activation part."""
    return [scale * torch.sigmoid(x) for x in logits]


def discriminator_input(image_pyramid, recon_pyramid) -> List[Tensor]:
    """train/utils.py:248-273 up to the discriminator call: detach_pyramid
    (:53-62) then concatenate_pyramids (:138-140)."""
    recon = [r.detach().clone() for r in recon_pyramid]
    return [torch.cat((x, y), 0) for x, y in zip(image_pyramid, recon)]


def combine_disparity(left: Tensor, right: Tensor, alpha: float = 20,
                      beta: float = 0.05) -> Tensor:
    """train/utils.py:199-245, the numpy original line by line."""
    left_disp = left.cpu().numpy()
    right_disp = right.cpu().numpy()
    mean_disp = (left_disp + right_disp) / 2
    _, height, width = mean_disp.shape
    x = np.linspace(0, 1, width)
    y = np.linspace(0, 1, height)
    xv, _ = np.meshgrid(x, y)
    left_mask = 1 - np.clip(alpha * (xv - beta), 0, 1)
    right_mask = np.fliplr(left_mask)
    mean_mask = 1 - (left_mask + right_mask)
    combined = (right_mask * left_disp) + (left_mask * right_disp) \
        + (mean_mask * mean_disp)
    return torch.from_numpy(combined)


def to_heatmap(x: Tensor, table: np.ndarray, inverse: bool = False) -> Tensor:
    """train/utils.py:177-196 with matplotlib's `Colormap.__call__` restated
    (matplotlib itself is absent from this image, so this part of the oracle is
    unpinned): for float input, xa = X * N in X's dtype, xa == N -> N - 1,
    xa < 0 -> under colour (= first entry for the stock maps), xa >= N -> over
    colour (= last entry), NaN -> bad colour (0, 0, 0, 0), else lut[int(xa)].
    `table`: the map's (N,3) fp64 RGB lookup table."""
    image = x.squeeze(0).cpu().numpy()
    image = 1 - image if inverse else image
    n = table.shape[0]
    xa = image * np.float32(n)                   # stays float32 like the input
    idx = np.where(np.isnan(xa), 0, xa).astype(np.int64)
    idx[xa == n] = n - 1
    idx[xa < 0] = 0
    idx[xa >= n] = n - 1
    idx[xa == n] = n - 1
    rgb = table[np.clip(idx, 0, n - 1)]
    rgb[np.isnan(xa)] = 0.0
    return torch.from_numpy(rgb).permute(2, 0, 1)
