"""torch-facing wrappers of the C ABI: argument checks, output allocation, stream hand-off and the
`torch.autograd.Function`s.  PyTorch is plumbing here (device memory, streams, autograd graph); every
arithmetic step of the hot path runs inside libgca_b200.so.
"""
import ctypes

import torch

from . import _lib
from ._lib import GCA_BF16, GCA_F32, ptr


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("gca_b200 ops run on CUDA tensors only (got a %s tensor); there is no CPU path" % t.device)


def _f32c(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def queue_dtype_code(queue):
    if queue.dtype == torch.float32:
        return GCA_F32
    if queue.dtype == torch.bfloat16:
        return GCA_BF16
    raise TypeError("queue must be float32 or bfloat16, got %s" % queue.dtype)


_WS = {}


def workspace(device, nbytes, tag="default"):
    """Per-(device, stream, tag) scratch buffer, grown on demand; never shrinks.  Keyed by the current stream because the
    kernels of one call keep state in it between launches: two streams must not share one."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, int(torch.cuda.current_stream(idx).cuda_stream), tag)
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.zeros(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _WS[key] = buf
    return buf


# --------------------------------------------------------------------------------------------- queue
def enqueue_(queue, keys, index, K_global=None, k_begin=0):
    """In-place ring-buffer enqueue (mem_moco.py:17-27).  Returns the new pointer (mem_moco.py:14-15).

    `queue` may be a K-shard holding global slots [k_begin, k_begin + len(queue)) of a ring of K_global."""
    _need_cuda(queue, keys)
    K_local, d = queue.shape
    K_global = K_local if K_global is None else int(K_global)
    if not queue.is_contiguous():
        raise ValueError("queue must be contiguous")
    keys = _f32c(keys.detach())
    N = keys.shape[0]
    if keys.shape[1] != d:
        raise ValueError("keys have %d features, queue has %d" % (keys.shape[1], d))
    _lib.call("gca_enqueue", ptr(queue), queue_dtype_code(queue), K_global, int(k_begin), int(k_begin) + K_local, d,
              ptr(keys), N, int(index), _stream(queue))
    return (int(index) + N) % K_global


# --------------------------------------------------------------------------------------------- InfoNCE
def infonce_workspace_bytes(B, K, d, queue_dtype, algo="auto"):
    return int(_lib.load().gca_infonce_workspace_bytes(B, K, d, queue_dtype, _lib.ALGO[algo]))


def infonce_forward(q, k, queue, T, algo="auto", want_grad=True, materialize=False):
    """Raw fused head: returns dict(loss, loss_rows, lse, pos, rank, dq_unit | None, logits | None)."""
    _need_cuda(q, k, queue)
    q, k = _f32c(q.detach()), _f32c(k.detach())
    B, d = q.shape
    K = queue.shape[0]
    dev = q.device
    qd = queue_dtype_code(queue)
    out = {
        "loss": torch.empty((), dtype=torch.float32, device=dev),
        "loss_rows": torch.empty(B, dtype=torch.float32, device=dev),
        "lse": torch.empty(B, dtype=torch.float32, device=dev),
        "pos": torch.empty(B, dtype=torch.float32, device=dev),
        "rank": torch.empty(B, dtype=torch.int32, device=dev),
        "hits": torch.empty(2, dtype=torch.int32, device=dev),
        "dq_unit": torch.empty(B, d, dtype=torch.float32, device=dev) if want_grad else None,
        "logits": torch.empty(B, K + 1, dtype=torch.float32, device=dev) if materialize else None,
    }
    ws = workspace(dev, infonce_workspace_bytes(B, K, d, qd, algo), "infonce")
    _lib.call("gca_infonce_fwd", ptr(q), ptr(k), ptr(queue), qd, B, K, d, 1.0 / T, _lib.ALGO[algo],
              ptr(out["loss"]), ptr(out["loss_rows"]), ptr(out["lse"]), ptr(out["pos"]), ptr(out["rank"]), ptr(out["hits"]),
              ptr(out["dq_unit"]), ptr(out["logits"]), ptr(ws), ws.numel(), _stream(q))
    return out


def moco_step_proj(zq, zk, queue, T, index, enqueue_keys=None, enqueue=True, algo="auto"):
    """gca_moco_step_proj: the head on UN-normalised projections zq, zk (the L2 normalisation of the projection head
    fused in), followed by the in-place enqueue of the normalised keys (or of `enqueue_keys`).  Returns a dict with
    loss, loss_rows, lse, pos, rank, hits, dz_unit (= d loss / d zq), k_hat (normalised keys) and the new pointer."""
    _need_cuda(zq, zk, queue)
    zq, zk = _f32c(zq.detach()), _f32c(zk.detach())
    B, d = zq.shape
    K = queue.shape[0]
    dev = zq.device
    qd = queue_dtype_code(queue)
    out = {
        "loss": torch.empty((), dtype=torch.float32, device=dev),
        "loss_rows": torch.empty(B, dtype=torch.float32, device=dev),
        "lse": torch.empty(B, dtype=torch.float32, device=dev),
        "pos": torch.empty(B, dtype=torch.float32, device=dev),
        "rank": torch.empty(B, dtype=torch.int32, device=dev),
        "hits": torch.empty(2, dtype=torch.int32, device=dev),
        "dz_unit": torch.empty(B, d, dtype=torch.float32, device=dev),
        "k_hat": torch.empty(B, d, dtype=torch.float32, device=dev),
    }
    keys = _f32c(enqueue_keys.detach()) if enqueue_keys is not None else None
    N = 0 if not enqueue else (keys.shape[0] if keys is not None else B)
    ws = workspace(dev, infonce_workspace_bytes(B, K, d, qd, algo), "infonce")
    _lib.call("gca_moco_step_proj", ptr(zq), ptr(zk), ptr(queue), qd, B, K, d, 1.0 / T, _lib.ALGO[algo],
              ptr(keys) if enqueue else None, N, int(index), None,
              ptr(out["loss"]), ptr(out["loss_rows"]), ptr(out["lse"]), ptr(out["pos"]), ptr(out["rank"]), ptr(out["hits"]),
              ptr(out["dz_unit"]), ptr(out["k_hat"]), ptr(ws), ws.numel(), _stream(zq))
    out["index"] = (int(index) + N) % K
    return out


class _InfoNCEFromProjections(torch.autograd.Function):
    """loss, ... = f(zq; zk, queue) with the projection head's Normalize(2) inside the kernels; the queue is updated in place
    by the same call (the enqueue rides in the last launch), so nothing is left for the caller to order."""

    @staticmethod
    def forward(ctx, zq, zk, queue, T, algo, index, all_k):
        o = moco_step_proj(zq, zk, queue, T, index, enqueue_keys=all_k, algo=algo)
        ctx.in_dtype = zq.dtype
        ctx.save_for_backward(o["dz_unit"])
        ctx.mark_non_differentiable(o["loss_rows"], o["lse"], o["pos"], o["rank"], o["k_hat"])
        return o["loss"], o["loss_rows"], o["lse"], o["pos"], o["rank"], o["k_hat"]      # (queue: a buffer, updated in place)

    @staticmethod
    def backward(ctx, g_loss, *_):
        (dz_unit,) = ctx.saved_tensors
        return (dz_unit * g_loss).to(ctx.in_dtype), None, None, None, None, None, None


def infonce_backward_recompute(q, k, queue, T, lse, grad_scale, algo="auto"):
    """Two-pass backward (gca_infonce_bwd): dq = grad_scale * d(sum_b loss_b)/dq against the given queue."""
    _need_cuda(q, k, queue, lse)
    q, k = _f32c(q.detach()), _f32c(k.detach())
    B, d = q.shape
    K = queue.shape[0]
    qd = queue_dtype_code(queue)
    dq = torch.empty(B, d, dtype=torch.float32, device=q.device)
    ws = workspace(q.device, infonce_workspace_bytes(B, K, d, qd, algo), "infonce")
    _lib.call("gca_infonce_bwd", ptr(q), ptr(k), ptr(queue), qd, B, K, d, 1.0 / T, _lib.ALGO[algo],
              ptr(_f32c(lse)), float(grad_scale), ptr(dq), ptr(ws), ws.numel(), _stream(q))
    return dq


class _InfoNCEFused(torch.autograd.Function):
    """loss (mean CE vs label 0), loss_rows, lse, pos, rank = f(q; k, queue).  Single pass: the unit gradient
    d loss / d q is produced by the forward kernels, so backward is one multiply and never
    re-reads the queue -- which the enqueue that follows the logits (mem_moco.py:82) is free to overwrite."""

    @staticmethod
    def forward(ctx, q, k, queue, T, algo):
        want = ctx.needs_input_grad[0]
        o = infonce_forward(q, k, queue, T, algo, want_grad=want)
        ctx.B = q.shape[0]
        ctx.in_dtype = q.dtype
        if want:
            ctx.save_for_backward(o["dq_unit"])
        ctx.mark_non_differentiable(o["lse"], o["pos"], o["rank"])
        return o["loss"], o["loss_rows"], o["lse"], o["pos"], o["rank"]

    @staticmethod
    def backward(ctx, g_loss, g_rows, *_):
        (dq_unit,) = ctx.saved_tensors
        dq = None
        if g_loss is not None:
            dq = dq_unit * g_loss
        if g_rows is not None:                       # d loss_rows[b] / d q_b = B * dq_unit[b]
            extra = dq_unit * (g_rows * float(ctx.B)).unsqueeze(1)
            dq = extra if dq is None else dq + extra
        if dq is not None and dq.dtype != ctx.in_dtype:
            dq = dq.to(ctx.in_dtype)
        return dq, None, None, None, None


def infonce_fused(q, k, queue, T, algo="auto"):
    return _InfoNCEFused.apply(q, k.detach(), queue, float(T), algo)


# --------------------------------------------------------------------------------------------- graph head
class _GraphCore(torch.autograd.Function):
    """y = GCN-aggregate(resample(hop-weight(softmax(gq . gk)))) for one layer (temporal_graph.py:161-239)."""

    @staticmethod
    def forward(ctx, gq, gk, support, u, alpha, max_hop, temperature, variants=None):
        _need_cuda(gq, gk, support, u)
        gq, gk, support, u = _f32c(gq), _f32c(gk), _f32c(support), _f32c(u)
        B, Cq, T = gq.shape[:3]
        S = gq[0, 0, 0].numel()
        C = support.shape[1]
        HW = support[0, 0, 0].numel()
        dev = gq.device
        sim = torch.empty(B, T, T, dtype=torch.float32, device=dev)
        adj = torch.empty_like(sim)
        s = torch.empty_like(sim)
        y = torch.empty_like(support)
        ws = workspace(dev, int(_lib.load().gca_graph_workspace_bytes(B, T)), "graph")
        opts = _graph_opts(variants)
        _lib.call("gca_graph_fwd_ex", ptr(gq), ptr(gk), Cq, S, ptr(support), C, HW, T, B, ptr(u), float(alpha), int(max_hop),
                  float(temperature), ctypes.byref(opts) if opts is not None else None, ptr(sim), ptr(adj), ptr(s), ptr(y),
                  ptr(ws), ws.numel(), _stream(gq))
        ctx.save_for_backward(gq, gk, support, sim, adj, s, u)
        ctx.cfg = (float(alpha), int(max_hop), float(temperature))
        ctx.variants = variants
        ctx.mark_non_differentiable(sim, adj, s)
        return y, sim, adj, s

    @staticmethod
    def backward(ctx, dy, *_):
        gq, gk, support, sim, adj, s, u = ctx.saved_tensors
        alpha, max_hop, temperature = ctx.cfg
        opts = _graph_opts(ctx.variants)
        dy = _f32c(dy)
        B, Cq, T = gq.shape[:3]
        S = gq[0, 0, 0].numel()
        C = support.shape[1]
        HW = support[0, 0, 0].numel()
        d_gq, d_gk, d_sup = torch.empty_like(gq), torch.empty_like(gk), torch.empty_like(support)
        ws = workspace(gq.device, int(_lib.load().gca_graph_workspace_bytes(B, T)), "graph")
        _lib.call("gca_graph_bwd_ex", ptr(gq), ptr(gk), Cq, S, ptr(support), C, HW, T, B, ptr(sim), ptr(adj), ptr(s), ptr(dy),
                  ptr(u), alpha, max_hop, temperature, ctypes.byref(opts) if opts is not None else None, ptr(d_gq), ptr(d_gk),
                  ptr(d_sup), ptr(ws), ws.numel(), _stream(gq))
        return d_gq, d_gk, d_sup, None, None, None, None, None


def _graph_opts(variants):
    """dict(threshold=tau, topk=k, edge_drop=p, symnorm=bool) -> GcaGraphOpts (None: the reference's arithmetic)."""
    if not variants:
        return None
    unknown = set(variants) - {"threshold", "topk", "edge_drop", "symnorm"}
    if unknown:
        raise ValueError("unknown graph variants %r" % sorted(unknown))
    o = _lib.GraphOpts(0, 0.0, 0, 0.0)
    if variants.get("threshold") is not None:
        o.flags |= _lib.GRAPH_THRESHOLD; o.tau = float(variants["threshold"])
    if variants.get("topk") is not None:
        o.flags |= _lib.GRAPH_TOPK; o.topk = int(variants["topk"])
    if variants.get("edge_drop") is not None:
        o.flags |= _lib.GRAPH_EDGE_DROP; o.p_drop = float(variants["edge_drop"])
    if variants.get("symnorm"):
        o.flags |= _lib.GRAPH_SYMNORM
    return o if o.flags else None


def graph_core(gq, gk, support, u, alpha=0.5, max_hop=3, temperature=1.0, variants=None):
    """gq, gk [B,Cq,T,...], support [B,C,T,...], u [B,T,T] -> (y like support, sim, adj, s).
    `variants`: default-OFF options with no reference counterpart (include/gca_b200.h GCA_GRAPH_*; parity unpinned):
    dict(threshold=tau, topk=k, edge_drop=p_drop, symnorm=True)."""
    return _GraphCore.apply(gq, gk, support, u, alpha, max_hop, temperature, variants)


# --------------------------------------------------------------------------------------------- SimSiam D
class _NegCos(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, z):
        _need_cuda(p, z)
        ctx.in_dtype = p.dtype
        ctx.shape = p.shape
        p2, z2 = _f32c(p.detach()).reshape(-1, p.shape[-1]), _f32c(z.detach()).reshape(-1, z.shape[-1])
        B, d = p2.shape
        loss = torch.empty((), dtype=torch.float32, device=p.device)
        want = ctx.needs_input_grad[0]
        dp = torch.empty_like(p2) if want else None
        ws = workspace(p.device, int(_lib.load().gca_negcos_workspace_bytes(B, d)), "negcos")
        _lib.call("gca_negcos_fwd_bwd", ptr(p2), ptr(z2), B, d, ptr(loss), None, ptr(dp), ptr(ws), ws.numel(), _stream(p))
        if want:
            ctx.save_for_backward(dp)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dp,) = ctx.saved_tensors
        return (dp * g).reshape(ctx.shape).to(ctx.in_dtype), None


def neg_cosine(p, z):
    """-mean cosine_similarity(p, stopgrad(z), dim=-1) (criterion.py:60)."""
    return _NegCos.apply(p, z.detach())


# --------------------------------------------------------------------------------------------- retrieval
class _BN1d(torch.autograd.Function):
    """BatchNorm1d (+ ReLU) on a [B, C] GEMM output: one launch forward, one backward (csrc/bn1d.cu)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, eps, momentum, relu, training):
        _need_cuda(x)
        x = _f32c(x)
        B, C = x.shape
        y = torch.empty_like(x)
        mean = torch.empty(C, dtype=torch.float32, device=x.device)
        invstd = torch.empty_like(mean)
        use_running = 0 if training else 1
        _lib.call("gca_bn1d_fwd", ptr(x), B, C, ptr(gamma), ptr(beta), float(eps), float(momentum), 1 if relu else 0,
                  use_running, ptr(running_mean), ptr(running_var), ptr(y), ptr(mean), ptr(invstd), _stream(x))
        ctx.save_for_backward(x, gamma, beta, mean, invstd)
        ctx.cfg = (1 if relu else 0, use_running)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, beta, mean, invstd = ctx.saved_tensors
        relu, use_running = ctx.cfg
        B, C = x.shape
        dy = _f32c(dy)
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dg = torch.empty(C, dtype=torch.float32, device=x.device) if gamma is not None else None
        db = torch.empty(C, dtype=torch.float32, device=x.device) if beta is not None else None
        _lib.call("gca_bn1d_bwd", ptr(x), ptr(dy), B, C, ptr(gamma), ptr(beta), ptr(mean), ptr(invstd), relu, use_running,
                  ptr(dx), ptr(dg), ptr(db), _stream(x))
        return dx, dg, db, None, None, None, None, None, None


def bn1d(x, bn, relu=False):
    """`relu(bn(x))` for an `nn.BatchNorm1d` module `bn` on a [B, C] tensor, with the module's own semantics: batch
    statistics + running-statistics update in training mode (or when it tracks none), running statistics in eval mode."""
    if x.dim() != 2:
        raise ValueError("bn1d expects [B, C], got %r" % (tuple(x.shape),))
    training = bn.training or bn.running_mean is None
    if training and x.shape[0] < 2:
        raise ValueError("Expected more than 1 value per channel when training, got input size %r" % (tuple(x.shape),))
    momentum = 0.0
    rm, rv = (bn.running_mean, bn.running_var) if bn.track_running_stats else (None, None)
    if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
        momentum = 1.0 / float(bn.num_batches_tracked) if bn.momentum is None else bn.momentum
    if not bn.training:
        momentum = 0.0
    return _BN1d.apply(x, bn.weight, bn.bias, rm, rv, bn.eps, momentum, relu, training)


def cosine_topk(queries, gallery, k, normalize=True):
    """Indices [Nq, k] (int32) and cosine similarities of the k nearest gallery rows of every query."""
    _need_cuda(queries, gallery)
    queries, gallery = _f32c(queries), _f32c(gallery)
    Nq, d = queries.shape
    Ng = gallery.shape[0]
    idx = torch.empty(Nq, k, dtype=torch.int32, device=queries.device)
    val = torch.empty(Nq, k, dtype=torch.float32, device=queries.device)
    ws = workspace(queries.device, int(_lib.load().gca_sim_topk_workspace_bytes(Nq, Ng, d, k)), "topk")
    _lib.call("gca_sim_topk", ptr(queries), ptr(gallery), Nq, Ng, d, k, 1 if normalize else 0, ptr(idx), ptr(val),
              ptr(ws), ws.numel(), _stream(queries))
    return idx, val


# --------------------------------------------------------------------------------------------- NPID instance bank
class _BankLogits(torch.autograd.Function):
    """logits[b, j] = <bank[idx[b, j]], x[b]> / T  (mem_bank.py:67-73, 29-39) without the [B, K+1, d] gather; the gradient
    reaches x only (the bank is a buffer)."""

    @staticmethod
    def forward(ctx, x, bank, idx, T):
        xc = _f32c(x.detach())
        B, d = xc.shape
        K1 = idx.shape[1]
        logits = torch.empty(B, K1, dtype=torch.float32, device=xc.device)
        _lib.call("gca_bank_logits", ptr(xc), ptr(bank), ptr(idx), B, K1, d, bank.shape[0], 1.0 / T, ptr(logits), _stream(xc))
        ctx.save_for_backward(bank, idx)
        ctx.T, ctx.shape = T, (B, K1, d)
        # the backward gathers the rows again: it must see the bank as it was here, and the caller updates the bank in place
        # right after the logits (mem_bank.py:81-85) -- so the rows' version is checked like any saved tensor
        return logits

    @staticmethod
    def backward(ctx, g):
        bank, idx = ctx.saved_tensors
        B, K1, d = ctx.shape
        g = _f32c(g)
        dx = torch.empty(B, d, dtype=torch.float32, device=g.device)
        ws = workspace(g.device, int(_lib.load().gca_bank_dx_workspace_bytes(B, d)), "bank_dx")
        _lib.call("gca_bank_dx", ptr(g), ptr(bank), ptr(idx), B, K1, d, bank.shape[0], 1.0 / ctx.T, ptr(dx), ptr(ws), ws.numel(),
                  _stream(g))
        return dx, None, None, None


def bank_logits(x, bank, idx, T):
    """x [B, d], bank [n_data, d] fp32 contiguous, idx [B, K+1] int64 -> logits [B, K+1] (differentiable w.r.t. x)."""
    _need_cuda(x, bank, idx)
    if bank.dtype != torch.float32 or not bank.is_contiguous():
        raise ValueError("the bank must be a contiguous fp32 tensor")
    idx = idx.contiguous()
    if idx.dtype != torch.int64:
        idx = idx.long()
    return _BankLogits.apply(x, bank, idx, float(T))


def bank_update_(bank, x, y, m):
    """In place: bank[y] <- normalize(m * bank[y] + (1 - m) * x)  (mem_bank.py:15-27)."""
    _need_cuda(bank, x, y)
    if bank.dtype != torch.float32 or not bank.is_contiguous():
        raise ValueError("the bank must be a contiguous fp32 tensor")
    x = _f32c(x.detach())
    y = y.reshape(-1).contiguous()
    if y.dtype != torch.int64:
        y = y.long()
    N, d = x.shape
    if y.numel() != N or d != bank.shape[1]:
        raise ValueError("x [N, d] and y [N] must match the bank's width")
    ws = workspace(bank.device, int(_lib.load().gca_bank_update_workspace_bytes(N, d)), "bank_update")
    # (1 - m) is a Python double in the reference and is rounded to fp32 once, when it meets the tensor
    _lib.call("gca_bank_update", ptr(bank), ptr(x), ptr(y), N, d, bank.shape[0], float(m), float(1.0 - m), ptr(ws), ws.numel(),
              _stream(bank))
    return bank
