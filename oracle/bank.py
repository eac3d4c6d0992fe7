"""CPU restatement of the NPID instance bank (reference: lib/memory/mem_bank.py, lib/memory/alias_multinomial.py,
lib/memory/criterion.py:8-31).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Parity status: PINNED -- oracle/gen_golden_bank.py runs the reference's own RGBMem / CMCMem / AliasMethod / NCECriterion on
seeded CPU inputs and checks every function here against them; tests/golden/bank.npz holds the vectors."""
import torch


def alias_tables(probs):
    """Vose alias tables as the reference builds them (alias_multinomial.py:8-42): returns (prob [K] fp32, alias [K] int64).
    The order in which outcomes are popped decides the tables, so the two work lists are handled exactly as upstream."""
    probs = probs.clone().float()
    if probs.sum() > 1:
        probs.div_(probs.sum())
    K = len(probs)
    prob = torch.zeros(K)
    alias = torch.zeros(K, dtype=torch.long)
    smaller, larger = [], []
    for kk in range(K):
        prob[kk] = K * probs[kk]
        (smaller if prob[kk] < 1.0 else larger).append(kk)
    while smaller and larger:
        small, large = smaller.pop(), larger.pop()
        alias[small] = large
        prob[large] = (prob[large] - 1.0) + prob[small]
        (smaller if prob[large] < 1.0 else larger).append(large)
    for last in smaller + larger:
        prob[last] = 1
    return prob, alias


def alias_pick(prob, alias, kk, b):
    """The arithmetic of AliasMethod.draw after its two random draws (alias_multinomial.py:56-65): kk = uniform outcome,
    b = bernoulli(prob[kk]); the sample is kk where b == 1, alias[kk] otherwise."""
    return kk * b.long() + alias.index_select(0, kk) * (1 - b).long()


def bank_logits(x, memory, idx, T):
    """mem_bank.py:68-73 + 29-39: logits[b, j] = <memory[idx[b, j]], x[b]> / T  (column 0 is the positive: idx[:, 0] = y)."""
    w = memory.index_select(0, idx.reshape(-1)).view(idx.shape[0], idx.shape[1], -1)
    return torch.bmm(w, x.unsqueeze(2)).squeeze(2) / T


def bank_grad_x(g_logits, memory, idx, T):
    """Gradient of bank_logits w.r.t. x (the bank is a buffer: no gradient reaches it): dx[b] = sum_j g[b, j] memory[idx[b, j]] / T."""
    w = memory.index_select(0, idx.reshape(-1)).view(idx.shape[0], idx.shape[1], -1)
    return torch.bmm(g_logits.unsqueeze(1), w).squeeze(1) / T


def bank_update(memory, x, y, m):
    """mem_bank.py:15-27, in place: rows y of the bank become normalize(m * row + (1 - m) * x); every occurrence of a
    duplicated index is computed from the OLD row and the last occurrence wins (index_copy_ on the CPU)."""
    w = memory.index_select(0, y.view(-1))
    w = w * m + x * (1 - m)
    w = w / w.norm(2, dim=1, keepdim=True).clamp_min(1e-12)
    # index_copy_ with a duplicated index is sequential on one CPU thread (how the fixture was generated: the last occurrence
    # wins) but unordered once ATen parallelises it; restate the sequential order explicitly
    yy = y.view(-1).tolist()
    for n, r in enumerate(yy):
        memory[r] = w[n]
    return memory


def nce_criterion(x, n_data):
    """criterion.py:14-31 (eps = 1e-7): x [B, m+1], column 0 the positive."""
    eps = 1e-7
    bsz, m = x.shape[0], x.shape[1] - 1
    Pn = 1.0 / float(n_data)
    p_pos = x[:, 0]
    log_d1 = (p_pos / (p_pos + (m * Pn + eps))).log()
    p_neg = x[:, 1:]
    log_d0 = (torch.full_like(p_neg, m * Pn) / (p_neg + (m * Pn + eps))).log()
    return -(log_d1.sum(0) + log_d0.reshape(-1, 1).sum(0)) / bsz
