"""Drop-in mirror of the reference's `lib.ops` (lib/ops/__init__.py:1, lib/ops/module_wrappers/__init__.py:1)."""
from .graph_head import TemporalGraphAug, GCN                     # noqa: F401
from .factory import build_aug_block, get_agg, TemporalAggreModel  # noqa: F401
from .mlp import ProjectionMLP, PredictionMLP                       # noqa: F401
