"""`build_aug_block` / `get_agg` with the reference's signatures (lib/ops/build.py:5-32).

The upstream `build_aug_block` cannot run (for/else bound to the inner loop, leaked loop variable, only the last
module wrapped: SURVEY.md R4).  This one does what it means to: every named sub-module `m` of the backbone is
replaced by `Sequential(TemporalGraphAug(in_channels_of(m)), m)`.
"""
import torch
import torch.nn as nn

from .graph_head import TemporalGraphAug


class TemporalAggreModel(nn.Module):
    """Segment pooling for 2D backbones (lib/ops/pooling_opts/basic_ops_wrap.py:4-27): mean / max over the segment
    axis.  Not on the 3D hot path (visual_wrappers.py:96-97 returns before it); kept for interface parity."""

    def __init__(self, pooling='avg', model_type='2D'):
        super(TemporalAggreModel, self).__init__()
        if pooling not in ('avg', 'max'):
            raise NotImplementedError("pooling %r" % (pooling,))
        self.pooling, self.model_type = pooling, model_type
        self.dim = 1 if model_type == '2D' else 2

    def forward(self, x):
        return torch.mean(x, dim=self.dim) if self.pooling == 'avg' else torch.max(x, dim=self.dim)


def get_agg(agg_fun='avg', model_type='2D'):
    return TemporalAggreModel(pooling=agg_fun, model_type='2D')      # upstream ignores model_type (build.py:6)


def _input_channels(module):
    if hasattr(module, 'in_channels'):
        return module.in_channels
    for m in module.modules():
        if isinstance(m, (nn.Conv3d, nn.Conv2d)):
            return m.in_channels
    raise ValueError("cannot infer the input channels of %s" % type(module).__name__)


def build_aug_block(base_model, module_name_list, n_segments):
    for name in module_name_list:
        parts = name.split('.')
        parent = base_model
        for p in parts[:-1]:
            parent = getattr(parent, p)
        leaf = parts[-1]
        target = parent[int(leaf)] if (leaf.isdigit() and isinstance(parent, (nn.Sequential, nn.ModuleList))) \
            else getattr(parent, leaf)
        wrapped = nn.Sequential(TemporalGraphAug(in_channels=_input_channels(target)), target)
        if leaf.isdigit() and isinstance(parent, (nn.Sequential, nn.ModuleList)):
            parent[int(leaf)] = wrapped
        else:
            setattr(parent, leaf, wrapped)
    return base_model
