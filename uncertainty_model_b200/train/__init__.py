from . import loss, sparsification, utils  # noqa: F401
from .loss import (ConsistencyLoss, GeneratorLoss, PerceptualLoss,  # noqa: F401
                   ReprojectionErrorLoss, SmoothnessLoss,
                   TukraUncertaintyLoss, WeightedSSIMLoss)
