"""B200-native stereo-uncertainty loss path.

Drop-in for the hot path of Probabilistic-Surgical-Vision/uncertainty-model:
`uncertainty_model_b200.train.{loss,utils,sparsification}` mirror the public
names of the reference's `train/loss.py`, `train/utils.py` (pyramid + warp
helpers) and `train/sparsification.py`; the arithmetic runs in hand-written
sm_100a CUDA kernels behind the C ABI of `include/usl.h` (libusl.so).
"""
from . import _lib  # noqa: F401

__all__ = ['train', 'functional', 'distributed']
