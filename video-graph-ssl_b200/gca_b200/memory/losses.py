"""Contrastive criteria behind the reference's interfaces (lib/memory/criterion.py:34-62)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import functional as GF
from .moco_queue import FusedLogits


class NCESoftmaxLoss(nn.Module):
    """Softmax cross-entropy against label 0 (criterion.py:34-45).  Given the `FusedLogits` handle of the fused
    head the loss has already been reduced inside the kernel; a real logits tensor takes the library path."""

    def __init__(self):
        super(NCESoftmaxLoss, self).__init__()

    def forward(self, x):
        if isinstance(x, FusedLogits):
            return x.loss
        label = torch.zeros(x.shape[0], dtype=torch.long, device=x.device)
        return F.cross_entropy(x, label)


class NCECriterion(nn.Module):
    """NCE loss of the instance-bank mode (criterion.py:8-31; eps = 1e-7): x [bsz, m+1] with the positive in column 0.
    Elementwise work on the bank logits, left to the framework."""

    def __init__(self, n_data):
        super(NCECriterion, self).__init__()
        self.n_data = n_data

    def forward(self, x):
        eps = 1e-7
        bsz = x.shape[0]
        m = x.size(1) - 1
        Pn = 1 / float(self.n_data)                                   # noise distribution
        P_pos = x.select(1, 0)
        log_D1 = torch.div(P_pos, P_pos.add(m * Pn + eps)).log_()
        P_neg = x.narrow(1, 1, m)
        log_D0 = torch.div(P_neg.clone().fill_(m * Pn), P_neg.add(m * Pn + eps)).log_()
        return - (log_D1.sum(0) + log_D0.view(-1, 1).sum(0)) / bsz


class D(nn.Module):
    """SimSiam negative cosine (criterion.py:47-62; duplicate at lib/modeling/graph_wrappers.py:93-108)."""

    def __init__(self, fun_type='v2'):
        super(D, self).__init__()
        if fun_type not in ('v1', 'v2'):
            raise NotImplementedError('Unknown type in simsiam D!')
        self.fun_type = fun_type

    def forward(self, logit, feat):
        # v1 (normalise both, dot, mean) and v2 (cosine_similarity) are the same function; one fused kernel
        return GF.neg_cosine(logit, feat.detach())
