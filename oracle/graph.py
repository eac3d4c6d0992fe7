"""Oracle: temporal clip-graph head (TemporalGraphAug).  TEST INFRASTRUCTURE ONLY.

Restates lib/ops/module_wrappers/temporal_graph.py: hop graph :25-36, similarity adjacency
:150-178, hop weighting :204-210, relaxed-Bernoulli re-sampling :187-192 (formula from
torch/distributions/relaxed_bernoulli.py `LogitRelaxedBernoulli.rsample` + SigmoidTransform),
GCN aggregation :56-64, forward :227-239.  The uniform noise `u` is an explicit input so the
result is a pure function (the reference draws it with `torch.rand` from the global generator).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

EPS32 = float(torch.finfo(torch.float32).eps)


def hop_distance(T: int, max_hop: int) -> np.ndarray:
    """[T, T] int32 hop distance on the chain graph with self loops; -1 where > max_hop.

    temporal_graph.py:25-36 builds it from powers of the chain adjacency; on a chain the hop
    distance is |i - j| (the reference stores +inf where unreachable within max_hop).
    """
    i = np.arange(T)
    hop = np.abs(i[:, None] - i[None, :]).astype(np.int32)
    hop[hop > max_hop] = -1
    return hop


def hop_weights(max_hop: int, alpha: float) -> np.ndarray:
    """theta(h) = e^-h / (1 + e^-2h) + alpha for h = 0..max_hop, computed in double like the
    reference's python lambda (temporal_graph.py:206)."""
    return np.array([math.exp(-h) / (1 + math.exp(-h) ** 2) + alpha for h in range(max_hop + 1)],
                    dtype=np.float64)


def _project(x, w, sub_sample, max_pool):
    g = F.conv3d(x, w)                                   # 1x1x1, bias-free (:119-122)
    if sub_sample:                                       # :127-129
        g = F.max_pool3d(g, (1, 2, 2)) if max_pool else F.avg_pool3d(g, (1, 2, 2))
    return g


def edge_weight_matrix(T, max_hop, alpha, dtype=torch.float32):
    hop = hop_distance(T, max_hop)
    th = hop_weights(max_hop, alpha)
    w = np.where(hop >= 0, th[np.clip(hop, 0, max_hop)], 0.0)
    return torch.from_numpy(w).to(dtype)


def graph_core(gq, gk, support, u, alpha=0.5, max_hop=3, temperature=1.0):
    """The part of the head the CUDA kernel owns.

    gq, gk : [B, Cq, T, S]   projections in conv-output layout (S = H'*W')
    support: [B, C, T, HW]   GCN 1x1x1 conv output
    u      : [B, T, T]       uniforms in [0, 1)
    returns (y[B,C,T,HW], sim[B,T,T], adj[B,T,T], s[B,T,T])
    """
    B, Cq, T, S = gq.shape
    eps = float(torch.finfo(gq.dtype).eps)
    # :161-167  flatten per time step in (c', h', w') order, dot products over that axis
    Gq = gq.transpose(2, 1).contiguous().view(B, T, -1)
    Gk = gk.transpose(2, 1).contiguous().view(B, T, -1)
    sim = F.softmax(torch.matmul(Gq, Gk.permute(0, 2, 1)), dim=-1)                 # :167, :176
    adj = sim * edge_weight_matrix(T, max_hop, alpha, sim.dtype)[None]             # :204-210
    p = adj.clamp(min=eps, max=1 - eps)                                            # clamp_probs
    uc = u.clamp(min=eps, max=1 - eps)
    z = (uc.log() - (-uc).log1p() + p.log() - (-p).log1p()) / temperature          # rsample
    s = torch.sigmoid(z)                                                           # SigmoidTransform
    y = torch.einsum('bij,bcjs->bcis', s, support) + support                       # :59-62
    return y, sim, adj, s


def graph_forward(x, wq, wk, wg, u, alpha=0.5, max_hop=3, temperature=1.0,
                  sub_sample=True, max_pool=True):
    """Whole `TemporalGraphAug.forward` (one GCN layer, the shipped default) as a function.

    x [B,C,T,H,W]; wq, wk [C',C,1,1,1]; wg [C,C,1,1,1]; returns (y [B,C,T,H,W], sim, adj, s).
    """
    B, C, T, H, W = x.shape
    gq = _project(x, wq, sub_sample, max_pool)
    gk = _project(x, wk, sub_sample, max_pool)
    support = F.conv3d(x, wg)                                                      # :58
    y, sim, adj, s = graph_core(gq.reshape(B, gq.shape[1], T, -1), gk.reshape(B, gk.shape[1], T, -1),
                                support.reshape(B, support.shape[1], T, -1), u,
                                alpha, max_hop, temperature)
    return y.view(B, -1, T, H, W), sim, adj, s


def graph_core_backward(gq, gk, support, sim, adj, s, dy, alpha=0.5, max_hop=3, temperature=1.0):
    """Closed-form backward of `graph_core`: returns (d_gq, d_gk, d_support)."""
    T = gq.shape[2]
    eps = float(torch.finfo(gq.dtype).eps)
    d_support = torch.einsum('bij,bcis->bcjs', s, dy) + dy
    ds = torch.einsum('bcis,bcjs->bij', dy, support)
    p = adj.clamp(min=eps, max=1 - eps)
    inside = (adj >= eps) & (adj <= 1 - eps)              # clamp passes gradient on [min, max]
    d_adj = torch.where(inside, ds * s * (1 - s) / (temperature * p * (1 - p)), torch.zeros_like(ds))
    d_sim = d_adj * edge_weight_matrix(T, max_hop, alpha, sim.dtype)[None]
    d_logit = sim * (d_sim - (d_sim * sim).sum(dim=-1, keepdim=True))
    d_gq = torch.einsum('bij,bcjs->bcis', d_logit, gk)
    d_gk = torch.einsum('bij,bcis->bcjs', d_logit, gq)
    return d_gq, d_gk, d_support


def graph_forward_backward(x, wq, wk, wg, u, dy, **kw):
    """Autograd through `graph_forward`: returns (y, dx, dwq, dwk, dwg)."""
    x = x.detach().clone().requires_grad_(True)
    ws = [w.detach().clone().requires_grad_(True) for w in (wq, wk, wg)]
    y, _, _, _ = graph_forward(x, ws[0], ws[1], ws[2], u, **kw)
    grads = torch.autograd.grad(y, [x] + ws, dy)
    return (y.detach(),) + tuple(grads)


def graph_core_variants(gq, gk, support, u, alpha=0.5, max_hop=3, temperature=1.0, threshold=None, topk=None, edge_drop=None,
                        symnorm=False):
    """`graph_core` with the default-OFF variants of include/gca_b200.h (GCA_GRAPH_*).  PARITY UNPINNED: the reference has
    no thresholded / top-k edges, no hard edge drop and no D^-1/2 A D^-1/2 on its forward path (its
    compute_ppr, temporal_graph.py:212-219, is dead and broken), so this is OUR statement of north_star's wording, written in
    differentiable torch ops (autograd provides the backward the kernel is checked against).  Order: hop-weighted adjacency ->
    threshold -> row top-k (ties towards the lower column) -> relaxed-Bernoulli re-sample | hard edge drop [u >= p] ->
    symmetric normalisation -> aggregation + skip.  Removed edges become 0 exactly like hop-masked edges."""
    B, Cq, T, S = gq.shape
    eps = EPS32                     # the clamps are those of the fp32 path (clamp_probs), whatever precision this runs in
    Gq = gq.transpose(2, 1).contiguous().view(B, T, -1)
    Gk = gk.transpose(2, 1).contiguous().view(B, T, -1)
    sim = F.softmax(torch.matmul(Gq, Gk.permute(0, 2, 1)), dim=-1)
    adj = sim * edge_weight_matrix(T, max_hop, alpha, sim.dtype)[None]
    keep = torch.ones_like(adj, dtype=torch.bool)
    if threshold is not None:
        keep &= ~(adj < threshold)
    adj_m = torch.where(keep, adj, torch.zeros_like(adj))
    if topk is not None:
        a = adj_m.detach()
        col = torch.arange(T)
        better = (a[..., None, :] > a[..., :, None]) | ((a[..., None, :] == a[..., :, None]) & (col[None, :] < col[:, None]))
        rank = better.sum(-1)                                   # [B, T, T]: entries of the row that come first
        adj_m = torch.where(rank < topk, adj_m, torch.zeros_like(adj_m))
    if edge_drop is not None:
        s = torch.where(u >= edge_drop, adj_m, torch.zeros_like(adj_m))
    else:
        p = adj_m.clamp(min=eps, max=1 - eps)
        uc = u.clamp(min=eps, max=1 - eps)
        s = torch.sigmoid((uc.log() - (-uc).log1p() + p.log() - (-p).log1p()) / temperature)
    if symnorm:
        r = s.sum(-1).clamp(min=eps).rsqrt()
        s = s * r[:, :, None] * r[:, None, :]
    y = torch.einsum('bij,bcjs->bcis', s, support) + support
    return y, sim, adj_m, s
