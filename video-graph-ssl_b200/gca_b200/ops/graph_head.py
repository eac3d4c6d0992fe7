"""Temporal clip-graph augmentation on the B200 kernels, behind the reference's `TemporalGraphAug` interface
(lib/ops/module_wrappers/temporal_graph.py:66-239).

Parameter-holding submodules keep the upstream names (`g_q`, `g_k`, `gcns.0.conv`) so checkpoints interchange.
The learned 1x1x1 convolutions (+ spatial pooling) stay with cuDNN; everything between them -- per-video T x T
similarity, row softmax, hop mask and theta(hop) weights, relaxed-Bernoulli re-sampling, aggregation + skip, and
the whole backward of that chain -- is one kernel launch per direction (`gca_graph_fwd` / `gca_graph_bwd`).

The upstream constructor cannot run as shipped (`reset_parameter` vs `reset_parameters`); this one
performs the initialisation the upstream code intends (:131-147).
"""
import math

import torch
from torch import nn

from .. import functional as GF


class GCN(nn.Module):
    """Holder of the GCN 1x1x1 projection (temporal_graph.py:38-47); aggregation happens in the fused kernel."""

    def __init__(self, in_features, out_features=None, bias=False, skip=True):
        super(GCN, self).__init__()
        if not skip:
            raise NotImplementedError("the fused aggregation always adds the skip term (upstream default)")
        self.skip = skip
        self.in_features = in_features
        self.out_features = in_features if out_features is None else out_features
        self.conv = nn.Conv3d(in_features, self.out_features, kernel_size=(1, 1, 1), bias=bias)


class TemporalGraphAug(nn.Module):
    def __init__(self, in_channels, inter_channels=None, sub_sample=True, bias=False, bn_layer=False,
                 zero_init=False, max_pool=True, mask_frame=False, nei_size=None, alpah=0.5,
                 num_gcn_layers=1, temperature=1., max_hop=3, *, adjacency="dot", edge_threshold=None, edge_topk=None,
                 edge_drop=None, sym_norm=False, feature_mask=None):
        """The positional / upstream keywords are the reference's (temporal_graph.py:67-71).  The keyword-only ones are
        default-OFF variants with NO reference counterpart (north_star's wording; parity unpinned, checked
        against oracle.graph.graph_core_variants): adjacency="cosine" (unit-norm per-frame projections before the dot
        product), edge_threshold=tau / edge_topk=k (sparsify the hop-weighted adjacency), edge_drop=p (hard seeded edge drop
        [u >= p] instead of the relaxed Bernoulli), sym_norm (D^-1/2 A D^-1/2), feature_mask=p (seeded per-(video, channel)
        mask of the output features)."""
        super(TemporalGraphAug, self).__init__()
        if adjacency not in ("dot", "cosine"):
            raise ValueError("adjacency must be 'dot' or 'cosine'")
        self.adjacency, self.feature_mask = adjacency, feature_mask
        self.variants = {k: v for k, v in (("threshold", edge_threshold), ("topk", edge_topk), ("edge_drop", edge_drop),
                                           ("symnorm", True if sym_norm else None)) if v is not None}
        if num_gcn_layers != 1:
            # upstream builds layers >= 2 from the raw `inter_channels` argument (None by default) and cannot
            # construct them (:94-99); one layer is the shipped configuration
            raise NotImplementedError("num_gcn_layers != 1 is not supported")
        self.sub_sample, self.bias, self.bn_layer = sub_sample, bias, bn_layer
        self.zero_init, self.max_pool = zero_init, max_pool
        self.in_channels = in_channels
        self.mask_frame, self.nei_size = mask_frame, nei_size
        self.alpha = alpah                                   # sic: the upstream keyword is spelled 'alpah'
        self.num_gcn_layers, self.temperature, self.max_hop = num_gcn_layers, temperature, max_hop
        self.inter_channels = max(in_channels // 2, 1) if inter_channels is None else inter_channels

        # :94-96 -- the GCN width comes from the RAW argument, i.e. in_channels -> in_channels by default
        self.gcns = nn.ModuleList([GCN(in_channels, inter_channels)])

        def projection():
            conv = nn.Conv3d(in_channels, self.inter_channels, kernel_size=1, stride=1, padding=0, bias=bias)
            return conv, (nn.Sequential(conv, nn.BatchNorm3d(self.inter_channels)) if bn_layer else conv)

        cq, self.g_q = projection()
        ck, self.g_k = projection()
        self.reset_parameters(cq, ck)
        if sub_sample:
            pool = nn.MaxPool3d(kernel_size=(1, 2, 2)) if max_pool else nn.AvgPool3d(kernel_size=(1, 2, 2))
            self.g_q = nn.Sequential(self.g_q, pool)
            self.g_k = nn.Sequential(self.g_k, pool)

    def reset_parameters(self, m1, m2):
        """U(-1/sqrt(fan_in), 1/sqrt(fan_in)) or zeros (:131-147), drawn in the upstream order (m1 then m2)."""
        for m in (m1, m2):
            if self.zero_init:
                nn.init.constant_(m.weight, 0)
            else:
                bound = 1. / math.sqrt(m.in_channels * m.kernel_size[0] * m.kernel_size[1] * m.kernel_size[2])
                m.weight.data.uniform_(-bound, bound)
        if self.bias:
            for m in (m1, m2):
                if self.zero_init:
                    nn.init.constant_(m.bias, 0)
                else:
                    bound = 1. / math.sqrt(m.in_channels * m.kernel_size[0] * m.kernel_size[1] * m.kernel_size[2])
                    m.bias.data.uniform_(-bound, bound)

    def forward(self, x, return_graph=False):
        """x [B, C, T, H, W] -> y [B, C', T, H, W] (C' = C by default)."""
        B, _, T, H, W = x.shape
        if self.mask_frame:
            # Upstream cannot complete a forward with mask_frame=True (probed on the reference itself): its mask loop indexes
            # the BATCH axis with range(nei_size) (temporal_graph.py:169-174) -> IndexError when B < nei_size (default T);
            # otherwise batch element 0 is fully masked, its softmax rows are NaN and RelaxedBernoulli's argument check
            # raises ValueError (:187-190).  Same error behaviour here, before any kernel runs.
            nei = T if not self.nei_size else self.nei_size
            if B < nei:
                raise IndexError("index %d is out of bounds for dimension 0 with size %d" % (B, B))
            raise ValueError("Expected parameter probs of distribution LogitRelaxedBernoulli to satisfy the constraint "
                             "Interval(lower_bound=0.0, upper_bound=1.0): mask_frame=True masks whole rows (NaN after softmax)")
        g_q = self.g_q(x)
        g_k = self.g_k(x)
        support = self.gcns[0].conv(x)                               # :58
        # the one RNG draw of the upstream forward: rsample's torch.rand(adj.shape) from the global generator (:188-191)
        u = torch.rand(B, T, T, dtype=torch.float32, device=x.device)
        if self.adjacency == "cosine":                               # variant: per-frame unit-norm projections (autograd)
            g_q = g_q / g_q.flatten(3).pow(2).sum(dim=(1, 3), keepdim=True).sqrt().clamp_min(1e-8).unsqueeze(-1)
            g_k = g_k / g_k.flatten(3).pow(2).sum(dim=(1, 3), keepdim=True).sqrt().clamp_min(1e-8).unsqueeze(-1)
        y, sim, adj, s = GF.graph_core(g_q, g_k, support, u, self.alpha, self.max_hop, self.temperature,
                                       self.variants or None)
        y = y.view(B, -1, T, H, W).to(x.dtype)
        if self.feature_mask is not None:                            # variant: drawn AFTER u (the default RNG stream is untouched)
            keep = torch.rand(B, y.shape[1], device=x.device) >= self.feature_mask
            y = y * keep[:, :, None, None, None].to(y.dtype)
        if return_graph:
            return y, {"sim": sim, "adj": adj, "s": s, "u": u}
        return y
