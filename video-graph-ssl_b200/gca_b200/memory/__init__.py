"""Drop-in mirror of the reference's `lib.memory` (lib/memory/__init__.py:1, build.py:5-32)."""
from .moco_queue import RGBMoCo, CMCMoCo, FusedLogits                 # noqa: F401
from .mem_bank import RGBMem, CMCMem, AliasMethod                     # noqa: F401
from .losses import NCESoftmaxLoss, NCECriterion, D                   # noqa: F401
from .factory import create_contrast, create_criterion                # noqa: F401
