"""SimSiam projection / prediction MLPs behind the reference's interfaces (lib/modeling/project_head.py:36-76).

Same constructor arguments, same sub-module layout (`l1`, `l2`, `l3` are `nn.Sequential`s of `nn.Linear`, `nn.BatchNorm1d`,
`nn.ReLU`), hence the same `state_dict()` keys and the same initialisation draws as upstream; only `forward` differs: every
`Linear -> BatchNorm1d (-> ReLU)` block runs as one library GEMM followed by ONE fused statistics + normalise + ReLU launch
(csrc/bn1d.cu) instead of ATen's three, forward and backward alike.
"""
import torch.nn as nn
import torch.nn.functional as F

from .. import functional as GF


def _linear_bn(block, x):
    lin, bn = block[0], block[1]
    relu = len(block) > 2 and isinstance(block[2], nn.ReLU)
    return GF.bn1d(F.linear(x, lin.weight, lin.bias), bn, relu=relu)


class ProjectionMLP(nn.Module):
    def __init__(self, in_dim, hid_dim, out_dim):
        super(ProjectionMLP, self).__init__()
        self.l1 = nn.Sequential(nn.Linear(in_dim, hid_dim), nn.BatchNorm1d(hid_dim), nn.ReLU(inplace=True))
        self.l2 = nn.Sequential(nn.Linear(hid_dim, hid_dim), nn.BatchNorm1d(hid_dim), nn.ReLU(inplace=True))
        self.l3 = nn.Sequential(nn.Linear(hid_dim, out_dim), nn.BatchNorm1d(out_dim))

    def forward(self, x):
        return _linear_bn(self.l3, _linear_bn(self.l2, _linear_bn(self.l1, x)))


class PredictionMLP(nn.Module):
    def __init__(self, in_dim, hid_dim, out_dim):
        super(PredictionMLP, self).__init__()
        self.l1 = nn.Sequential(nn.Linear(in_dim, hid_dim), nn.BatchNorm1d(hid_dim), nn.ReLU(inplace=True))
        self.l2 = nn.Linear(hid_dim, out_dim)

    def forward(self, x):
        return self.l2(_linear_bn(self.l1, x))
