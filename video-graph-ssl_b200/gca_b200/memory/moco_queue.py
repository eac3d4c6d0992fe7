"""MoCo memory queue on the B200 kernels, behind the reference's `RGBMoCo` / `CMCMoCo` interface
(lib/memory/mem_moco.py:52-142).

Same constructor, same `forward(q, k, q_jig=None, all_k=None) -> (logits, labels)` call, same `memory` buffer /
`K` / `T` / `index` attributes (train_video_contrast_dis.py:105-121, 239, 278, 411).  What changes underneath:
  * the queue is read in place (no 33 MB `memory.clone()` per step, mem_moco.py:72) -- stream order puts the
    logits kernel before this step's enqueue, and the gradient is produced in the same pass, so nothing ever
    re-reads rows the enqueue has overwritten;
  * the [B, K+1] logits are never materialised: `forward` returns a `FusedLogits` handle that carries the loss,
    the per-row rank of the positive and exactly the attributes the unmodified callers touch (`.shape[0]`,
    `.detach()`, `.topk()` for `accuracy`, and `NCESoftmaxLoss.forward`);  `materialize=True` returns a real
    tensor for parity checks;
  * `queue_dtype='bf16'` stores the queue in bfloat16 and runs the tcgen05 kernel (d == 128).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import functional as GF


class FusedLogits(object):
    """Stand-in for the [B, K+1] logits tensor of `RGBMoCo.forward` (mem_moco.py:45-46).

    `loss` is the mean cross-entropy against label 0 (criterion.py:44), autograd-connected to q;
    `rank[b]` counts the negatives strictly greater than the positive of row b.
    """

    def __init__(self, loss, loss_rows, lse, pos, rank, B, K):
        self.loss, self.loss_rows, self.lse, self.pos, self.rank = loss, loss_rows, lse, pos, rank
        self.shape = torch.Size((B, K + 1))
        self.device, self.dtype = loss.device, loss.dtype

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def dim(self):
        return 2

    def detach(self):
        return FusedLogits(self.loss.detach(), self.loss_rows.detach(), self.lse, self.pos, self.rank,
                           self.shape[0], self.shape[1] - 1)

    def topk(self, k, dim=1, largest=True, sorted=True):
        """What `accuracy` needs (lib/evaluation/metric.py:51): pred[b, r] == 0 iff the positive (column 0) is the
        r-th largest logit of row b.  Other entries are a non-zero filler; values are not reconstructed."""
        if dim not in (1, -1) or not largest:
            raise NotImplementedError("FusedLogits.topk supports dim=1, largest=True")
        r = torch.arange(k, device=self.rank.device).unsqueeze(0)
        pred = torch.where(r == self.rank.unsqueeze(1).long(), torch.zeros_like(r), r + 1)
        return torch.zeros(pred.shape, dtype=torch.float32, device=pred.device), pred


class BaseMoCo(nn.Module):
    """mem_moco.py:6-15: K slots, temperature T, python-int ring pointer `index` (not in state_dict, as upstream)."""

    def __init__(self, K=65536, T=0.07, queue_dtype="fp32", algo="auto", materialize=False):
        super(BaseMoCo, self).__init__()
        self.K = K
        self.T = T
        self.index = 0
        if queue_dtype not in ("fp32", "bf16"):
            raise ValueError("queue_dtype must be 'fp32' or 'bf16'")
        self.queue_dtype = queue_dtype
        self.algo = algo
        self.materialize = materialize

    def _update_pointer(self, bsz):
        self.index = (self.index + bsz) % self.K

    def _update_memory(self, k, queue):
        """mem_moco.py:17-27 as one in-place vectorised kernel."""
        with torch.no_grad():
            GF.enqueue_(queue, k, self.index)

    def _push(self, pairs):
        """Enqueue every (keys, queue) pair at the current pointer, then advance it once (mem_moco.py:81-83)."""
        n = pairs[0][0].size(0)
        for keys, queue in pairs:
            self._update_memory(keys, queue)
        self._update_pointer(n)

    @staticmethod
    def _labels(q):
        return torch.zeros(q.size(0), dtype=torch.long, device=q.device)      # mem_moco.py:78

    def _make_queue(self, n_dim):
        q = F.normalize(torch.randn(self.K, n_dim))          # same RNG draw and normalisation as mem_moco.py:57-58
        return q.to(torch.bfloat16) if self.queue_dtype == "bf16" else q

    def _head(self, q, k, queue):
        """Fused logits + loss (+ unit gradient) of one (q, k, queue) triple."""
        if self.materialize:
            o = GF.infonce_forward(q, k, queue, self.T, algo="ffma", want_grad=False, materialize=True)
            return _MaterializedLogits.apply(q, k.detach(), queue, o["logits"], self.T)
        loss, loss_rows, lse, pos, rank = GF.infonce_fused(q, k, queue, self.T, self.algo)
        return FusedLogits(loss, loss_rows, lse, pos, rank, q.shape[0], queue.shape[0])

    # checkpoints keep the upstream format: fp32 `memory` under the same key (train_video_contrast_dis.py:278)
    def _save_to_state_dict(self, destination, prefix, keep_vars):
        super(BaseMoCo, self)._save_to_state_dict(destination, prefix, keep_vars)
        for name, buf in self._buffers.items():
            if buf is not None and buf.dtype == torch.bfloat16:
                destination[prefix + name] = buf.float()

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        for name, buf in self._buffers.items():
            key = prefix + name
            if buf is not None and key in state_dict and state_dict[key].dtype != buf.dtype:
                state_dict[key] = state_dict[key].to(buf.dtype)
        super(BaseMoCo, self)._load_from_state_dict(state_dict, prefix, *args, **kwargs)


class _MaterializedLogits(torch.autograd.Function):
    """Parity / debugging path: real [B, K+1] logits from the kernel; the backward uses the queue as it was when
    the logits were taken (a snapshot, like the reference's clone at mem_moco.py:72)."""

    @staticmethod
    def forward(ctx, q, k, queue, logits, T):
        ctx.save_for_backward(k, queue.detach().clone())
        ctx.T = T
        return logits

    @staticmethod
    def backward(ctx, g):
        k, snap = ctx.saved_tensors
        dq = (g[:, :1] * k + g[:, 1:] @ snap.float()) / ctx.T
        return dq, None, None, None, None


class RGBMoCo(BaseMoCo):
    """Single-modality MoCo cache (mem_moco.py:52-88)."""

    def __init__(self, n_dim, K=65536, T=0.07, queue_dtype="fp32", algo="auto", materialize=False):
        super(RGBMoCo, self).__init__(K, T, queue_dtype, algo, materialize)
        self.register_buffer('memory', self._make_queue(n_dim))

    def forward(self, q, k, q_jig=None, all_k=None):
        k = k.detach()
        # heads first (queue read in place), enqueue after: the order of mem_moco.py:73 then :82
        heads = [self._head(x, k, self.memory) for x in (q, q_jig) if x is not None]
        self._push([(k if all_k is None else all_k, self.memory)])
        return tuple(heads) + (self._labels(q),)


    def forward_from_projections(self, zq, zk, all_k=None):
        """`forward(normalize(zq), normalize(zk), all_k=all_k)` with the projection head's trailing `Normalize(2)`
        (lib/modeling/project_head.py:4-10, 22-28) done inside the kernels: zq / zk are the outputs of the head's last
        Linear, the gradient flows back to zq through the normalisation, and the normalised keys (or `all_k`, already
        normalised rows gathered from every rank) are enqueued by the same call.  bf16 queue with n_dim == 128 only.
        Returns (FusedLogits, labels, k_hat) -- k_hat are the normalised keys, e.g. for the key all-gather."""
        loss, loss_rows, lse, pos, rank, k_hat = GF._InfoNCEFromProjections.apply(
            zq, zk.detach(), self.memory, self.T, self.algo, self.index, None if all_k is None else all_k.detach())
        n = zq.shape[0] if all_k is None else all_k.shape[0]
        self._update_pointer(n)
        return FusedLogits(loss, loss_rows, lse, pos, rank, zq.shape[0], self.K), self._labels(zq), k_hat


class CMCMoCo(BaseMoCo):
    """Two-modality variant (mem_moco.py:91-142): two queues, cross-modal positives, same kernels."""

    def __init__(self, n_dim, K=65536, T=0.07, queue_dtype="fp32", algo="auto", materialize=False):
        super(CMCMoCo, self).__init__(K, T, queue_dtype, algo, materialize)
        self.register_buffer('memory_1', self._make_queue(n_dim))
        self.register_buffer('memory_2', self._make_queue(n_dim))

    def forward(self, q1, k1, q2, k2, q1_jig=None, q2_jig=None, all_k1=None, all_k2=None):
        k1, k2 = k1.detach(), k2.detach()
        # modality 1 queries score against modality 2 keys/queue and vice versa (mem_moco.py:120-121)
        anchors = [(q1, k2, self.memory_2), (q2, k1, self.memory_1)]
        if q1_jig is not None and q2_jig is not None:
            anchors += [(q1_jig, k2, self.memory_2), (q2_jig, k1, self.memory_1)]
        heads = [self._head(*a) for a in anchors]
        new1 = k1 if all_k1 is None else all_k1
        new2 = k2 if all_k2 is None else all_k2
        if new1.size(0) != new2.size(0):
            raise AssertionError("both modalities must enqueue the same number of keys")
        self._push([(new1, self.memory_1), (new2, self.memory_2)])
        return tuple(heads) + (self._labels(q1),)
