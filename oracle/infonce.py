"""Oracle: MoCo logits + InfoNCE loss / gradient / top-k.  TEST INFRASTRUCTURE ONLY.

Restates lib/memory/mem_moco.py:29-88 (logits, forward), lib/memory/criterion.py:34-45
(`NCESoftmaxLoss`), lib/evaluation/metric.py:44-67 (`accuracy`) and the K-sharded combine of
SURVEY.md Appendix A.4.  All functions take CPU tensors; pass `.double()` inputs for an fp64 run.
"""
import math

import torch
import torch.nn.functional as F

from .ring import enqueue


def logits_full(q, k, queue, T):
    """[B, K+1] logits: column 0 = q_b.k_b / T, columns 1..K = q_b.queue_j / T.

    mem_moco.py:36-46: `bmm` for the positive, `mm(queue, q^T)^T` for the negatives, `cat`, `div`.
    `k` and `queue` carry no gradient (mem_moco.py:69, :72).
    """
    k = k.detach()
    queue = queue.detach().to(q.dtype)
    pos = (q * k).sum(dim=1, keepdim=True)
    neg = torch.mm(queue, q.t()).t()
    return torch.cat((pos, neg), dim=1) / T


def infonce_loss(logits):
    """criterion.py:40-45: mean cross-entropy against label 0 for every row."""
    return (torch.logsumexp(logits, dim=1) - logits[:, 0]).mean()


def infonce_grad_q(q, k, queue, T, grad_out=1.0):
    """Closed-form dL/dq of `infonce_loss(logits_full(...))`.

    dq_b = g/(T*B) * [ (p_b0 - 1) k_b + sum_j p_bj queue_j ],  p = softmax(logits_b).
    """
    with torch.no_grad():
        lg = logits_full(q, k, queue, T)
        p = torch.softmax(lg, dim=1)
        B = q.shape[0]
        acc = (p[:, :1] - 1.0) * k + p[:, 1:] @ queue.to(q.dtype)
        return acc * (grad_out / (T * B))


def positive_rank(logits):
    """Number of negatives STRICTLY greater than the positive, per row (int64).

    `accuracy` (metric.py:51) uses `topk`, whose order among exact ties is unspecified; counting
    strict wins makes the positive win ties.  top-k hit <=> rank < k.
    """
    return (logits[:, 1:] > logits[:, :1]).sum(dim=1)


def topk_accuracy(logits, topk=(1, 5)):
    """metric.py:44-67 for the single-label case with target 0 (mem_moco.py:78), in percent.

    The reference's `correct[:k].view(-1)` fails on torch>=1.7 (non-contiguous); `reshape` restates
    the intended arithmetic.
    """
    B = logits.shape[0]
    maxk = max(topk)
    _, pred = logits.topk(maxk, 1, True, True)
    correct = pred.t().eq(torch.zeros(1, B, dtype=pred.dtype))
    return [correct[:kk].reshape(-1).float().sum() * (100.0 / B) for kk in topk]


def reference_head_step(q, k, memory, index, T, all_k=None, topk=(1, 5)):
    """One full head step exactly as the trainer sequences it
    (tools/train_video_contrast_dis.py:411-428): RGBMoCo.forward (clone queue, logits, enqueue,
    pointer), NCESoftmaxLoss, backward to q, accuracy.  `q` must require grad.  `memory` is updated
    in place; returns (loss, dq, new_index, [top1, top5]).  This is the port that `bench.py` times
    as the CPU baseline.
    """
    snapshot = memory.clone().detach()                      # mem_moco.py:72
    lg = logits_full(q, k, snapshot, T)                     # mem_moco.py:73
    new_index = enqueue(memory, k if all_k is None else all_k, index)   # mem_moco.py:81-83
    label = torch.zeros(lg.shape[0], dtype=torch.long)     # criterion.py:43
    loss = F.cross_entropy(lg, label)                       # criterion.py:44
    (dq,) = torch.autograd.grad(loss, q)
    acc = topk_accuracy(lg.detach(), topk)
    return loss.detach(), dq, new_index, acc


def infonce_step(q, k, memory, index, T, all_k=None):
    """Closed-form version of `reference_head_step` (no autograd, fp32 or fp64 by input dtype).

    Returns dict(loss, loss_rows, lse, pos, rank, dq, index).  The gradient uses the queue as it
    was BEFORE this step's enqueue (mem_moco.py:72-73 run before :82).
    """
    with torch.no_grad():
        lg = logits_full(q, k, memory, T)
        lse = torch.logsumexp(lg, dim=1)
        out = {
            "pos": lg[:, 0].clone(),
            "lse": lse,
            "loss_rows": lse - lg[:, 0],
            "loss": (lse - lg[:, 0]).mean(),
            "rank": positive_rank(lg),
            "dq": infonce_grad_q(q, k, memory, T),
        }
        out["index"] = enqueue(memory, k if all_k is None else all_k, index)
    return out


def lse_partials(q, queue_shard, T, pos=None):
    """Per-row online-softmax partial over one K-shard: (m, s, cnt).

    m = max_j l_bj, s = sum_j exp(l_bj - m), cnt = #{j : l_bj > pos_b} over the shard's negatives
    only.  An empty shard gives (-inf, 0, 0).
    """
    with torch.no_grad():
        B = q.shape[0]
        if queue_shard.shape[0] == 0:
            return (torch.full((B,), -math.inf, dtype=q.dtype), torch.zeros(B, dtype=q.dtype),
                    torch.zeros(B, dtype=torch.int64))
        lg = (q @ queue_shard.to(q.dtype).t()) / T
        m = lg.max(dim=1).values
        s = torch.exp(lg - m[:, None]).sum(dim=1)
        cnt = (lg > pos[:, None]).sum(dim=1) if pos is not None else torch.zeros(B, dtype=torch.int64)
        return m, s, cnt


def merge_partials(ms, ss, pos):
    """Combine shard partials and the positive logit into the row log-sum-exp.

    M = max(max_r m_r, pos); S = sum_r s_r e^{m_r - M} + e^{pos - M}; lse = M + log S.
    `ms`, `ss`: [W, B].
    """
    M = torch.maximum(ms.max(dim=0).values, pos)
    S = (ss * torch.exp(ms - M[None, :])).sum(dim=0) + torch.exp(pos - M)
    return M + torch.log(S)


def sharded_infonce(q, k, queue, T, world):
    """Loss / gradient / rank computed shard-by-shard along K then merged: must equal the
    unsharded result (online-softmax associativity).  Returns dict(loss, lse, rank, dq)."""
    with torch.no_grad():
        K = queue.shape[0]
        assert K % world == 0
        Ks = K // world
        pos = (q * k).sum(dim=1) / T
        parts = [lse_partials(q, queue[r * Ks:(r + 1) * Ks], T, pos) for r in range(world)]
        ms = torch.stack([p[0] for p in parts])
        ss = torch.stack([p[1] for p in parts])
        rank = torch.stack([p[2] for p in parts]).sum(dim=0)
        lse = merge_partials(ms, ss, pos)
        B = q.shape[0]
        acc = (torch.exp(pos - lse) - 1.0)[:, None] * k
        for r in range(world):
            shard = queue[r * Ks:(r + 1) * Ks].to(q.dtype)
            p = torch.exp((q @ shard.t()) / T - lse[:, None])
            acc = acc + p @ shard
        return {"loss": (lse - pos).mean(), "lse": lse, "rank": rank, "dq": acc / (T * B)}


def head_from_projections(zq, zk, memory, index, T, all_k=None):
    """The head on the UN-normalised outputs of the projection head's last Linear: `Normalize(2)` of both
    (lib/modeling/project_head.py:4-10, applied at :22-28 as the last module of `ProjectHead.head`), then
    `infonce_step`.  Returns its dict plus `dz` = d loss / d zq through the normalisation
    (dz = (g - (g . q) q) / ||zq||, g = d loss / d q) and `k_hat`, the normalised keys that get enqueued."""
    q = F.normalize(zq, p=2, dim=1)
    k = F.normalize(zk, p=2, dim=1)
    out = infonce_step(q, k, memory, index, T, all_k=all_k)
    g = out["dq"].to(q.dtype)
    inv = 1.0 / zq.norm(dim=1, keepdim=True).clamp_min(1e-12)
    out["dz"] = (g - (g * q).sum(1, keepdim=True) * q) * inv
    out["k_hat"] = k
    return out
