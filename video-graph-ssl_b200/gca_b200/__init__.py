"""gca_b200 -- B200-native (sm_100a) contrastive hot path of GCA behind the reference's module interfaces.

Mirrors `lib.memory` (create_contrast / create_criterion / RGBMoCo / NCESoftmaxLoss / D) and `lib.ops`
(build_aug_block / get_agg / TemporalGraphAug) of ACMMM2021-Anonymous/video-graph-ssl; everything below those
interfaces runs in libgca_b200.so (hand-written CUDA, C ABI in include/gca_b200.h).  No CPU fallback.
"""
from . import _lib, functional                                   # noqa: F401
from .memory import (RGBMoCo, CMCMoCo, RGBMem, CMCMem, NCESoftmaxLoss, NCECriterion, D, FusedLogits,  # noqa: F401
                     create_contrast, create_criterion)
from .ops import TemporalGraphAug, build_aug_block, get_agg, ProjectionMLP, PredictionMLP       # noqa: F401

__all__ = ["RGBMoCo", "CMCMoCo", "RGBMem", "CMCMem", "NCESoftmaxLoss", "NCECriterion", "D", "FusedLogits", "create_contrast", "create_criterion",
           "TemporalGraphAug", "build_aug_block", "get_agg", "ProjectionMLP", "PredictionMLP", "functional"]
