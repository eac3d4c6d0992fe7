"""NPID instance bank behind the reference's interfaces (lib/memory/mem_bank.py: BaseMem, RGBMem, CMCMem;
lib/memory/alias_multinomial.py: AliasMethod).

Same constructors, buffers (`memory` / `memory_1`, `memory_2`) and forward signatures.  The sampled bank rows are scored by
gca_bank_logits without the reference's [bsz, K+1, n_dim] gather; the in-place momentum update is gca_bank_update.  The
negative sampler draws with the same two torch calls in the same order as upstream, so a seeded run draws the same indices.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import functional as GF


class AliasMethod(object):
    """Vose's alias method; tables built exactly like alias_multinomial.py:8-42 (the pop order decides them)."""

    def __init__(self, probs):
        probs = probs.clone().float()
        if probs.sum() > 1:
            probs.div_(probs.sum())
        K = len(probs)
        scaled = (probs * K).tolist()
        prob = [0.0] * K
        alias = [0] * K
        smaller, larger = [], []
        f32 = lambda v: float(torch.tensor(v, dtype=torch.float32))      # the reference keeps the tables in fp32
        for kk in range(K):
            prob[kk] = f32(scaled[kk])
            (smaller if prob[kk] < 1.0 else larger).append(kk)
        while smaller and larger:
            small, large = smaller.pop(), larger.pop()
            alias[small] = large
            prob[large] = f32(f32(prob[large] - 1.0) + prob[small])
            (smaller if prob[large] < 1.0 else larger).append(large)
        for last in smaller + larger:
            prob[last] = 1.0
        self.prob = torch.tensor(prob, dtype=torch.float32)
        self.alias = torch.tensor(alias, dtype=torch.long)

    def cuda(self):
        self.prob = self.prob.cuda()
        self.alias = self.alias.cuda()

    def to(self, device):
        self.prob, self.alias = self.prob.to(device), self.alias.to(device)
        return self

    def draw(self, N):
        """N samples (alias_multinomial.py:49-65): one uniform outcome and one Bernoulli per sample."""
        K = self.alias.size(0)
        kk = torch.zeros(N, dtype=torch.long, device=self.prob.device).random_(0, K)
        prob = self.prob.index_select(0, kk)
        alias = self.alias.index_select(0, kk)
        b = torch.bernoulli(prob)
        return kk.mul(b.long()) + alias.mul((1 - b).long())


class BaseMem(nn.Module):
    """mem_bank.py:7-39"""

    def __init__(self, K=65536, T=0.07, m=0.5):
        super(BaseMem, self).__init__()
        self.K = K
        self.T = T
        self.m = m

    def _sampler_to(self, device):
        if self.multinomial.prob.device != device:
            self.multinomial.to(device)

    def _draw_indices(self, y, bsz):
        self._sampler_to(y.device)
        idx = self.multinomial.draw(bsz * (self.K + 1)).view(bsz, -1)
        idx.select(1, 0).copy_(y.data)                                   # column 0 is the positive (mem_bank.py:64-66)
        return idx

    def _update_memory(self, memory, x, y):
        GF.bank_update_(memory, x, y, self.m)

    def _compute_logit(self, x, memory, idx):
        if not x.is_cuda:
            raise RuntimeError("gca_b200 memory banks run on CUDA tensors only; there is no CPU path")
        return GF.bank_logits(x, memory, idx, self.T)


class RGBMem(BaseMem):
    """Memory bank for a single modality (mem_bank.py:42-90)."""

    def __init__(self, n_dim, n_data, K=65536, T=0.07, m=0.5):
        super(RGBMem, self).__init__(K, T, m)
        self.multinomial = AliasMethod(torch.ones(n_data))
        self.register_buffer('memory', torch.randn(n_data, n_dim))
        self.memory = F.normalize(self.memory)

    def forward(self, x, y, x_jig=None, all_x=None, all_y=None, idx=None):
        """Same arguments as upstream; `idx` [bsz, K+1] (optional, tests) replaces the drawn indices."""
        bsz = x.size(0)
        if idx is None:
            idx = self._draw_indices(y, bsz)
        logits = self._compute_logit(x, self.memory, idx)
        logits_jig = self._compute_logit(x_jig, self.memory, idx) if x_jig is not None else None
        labels = torch.zeros(bsz, dtype=torch.long, device=x.device)
        # the update overwrites bank rows the backward pass gathers again: score against a snapshot only when a gradient is due
        if torch.is_grad_enabled() and (x.requires_grad or (x_jig is not None and x_jig.requires_grad)):
            self.memory = self.memory.clone()
        if (all_x is not None) and (all_y is not None):
            self._update_memory(self.memory, all_x, all_y)
        else:
            self._update_memory(self.memory, x, y)
        if x_jig is not None:
            return logits, logits_jig, labels
        return logits, labels


class CMCMem(BaseMem):
    """Memory bank for two modalities (mem_bank.py:93-160): head 1 scores x1 against modality 2's bank and vice versa."""

    def __init__(self, n_dim, n_data, K=65536, T=0.07, m=0.5):
        super(CMCMem, self).__init__(K, T, m)
        self.multinomial = AliasMethod(torch.ones(n_data))
        self.register_buffer('memory_1', torch.randn(n_data, n_dim))
        self.register_buffer('memory_2', torch.randn(n_data, n_dim))
        self.memory_1 = F.normalize(self.memory_1)
        self.memory_2 = F.normalize(self.memory_2)

    def forward(self, x1, x2, y, x1_jig=None, x2_jig=None, all_x1=None, all_x2=None, all_y=None, idx=None):
        bsz = x1.size(0)
        if idx is None:
            idx = self._draw_indices(y, bsz)
        logits1 = self._compute_logit(x1, self.memory_2, idx)
        logits2 = self._compute_logit(x2, self.memory_1, idx)
        jig = (x1_jig is not None) and (x2_jig is not None)
        if jig:
            logits1_jig = self._compute_logit(x1_jig, self.memory_2, idx)
            logits2_jig = self._compute_logit(x2_jig, self.memory_1, idx)
        labels = torch.zeros(bsz, dtype=torch.long, device=x1.device)
        if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (x1, x2, x1_jig, x2_jig)):
            self.memory_1, self.memory_2 = self.memory_1.clone(), self.memory_2.clone()
        if (all_x1 is not None) and (all_x2 is not None) and (all_y is not None):
            self._update_memory(self.memory_1, all_x1, all_y)
            self._update_memory(self.memory_2, all_x2, all_y)
        else:
            self._update_memory(self.memory_1, x1, y)
            self._update_memory(self.memory_2, x2, y)
        if jig:
            return logits1, logits2, logits1_jig, logits2_jig, labels
        return logits1, logits2, labels
