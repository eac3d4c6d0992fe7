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
        rows, negatives = x.shape[0], x.shape[1] - 1
        noise = negatives / float(self.n_data)                        # m * Pn: expected noise mass against one data sample
        shift = noise + 1e-7                                          # (eps of the reference, added in double like upstream)
        log_d1 = (x[:, 0] / (x[:, 0] + shift)).log()                  # positives: P(data | x)
        log_d0 = (noise / (x[:, 1:] + shift)).log()                   # negatives: P(noise | x)
        return -(log_d1.sum() + log_d0.sum()).reshape(1) / rows       # shape [1], as upstream returns it


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
