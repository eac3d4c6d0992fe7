#!/usr/bin/env python
"""Runs the c3-fmap graph head forward+backward a few times (for an ncu launch list / per-kernel times)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-graph-ssl_b200"))
from gca_b200 import functional as GF
torch.manual_seed(0)
B, C, T, H = 128, 192, 8, 14
dev = "cuda"
gq = torch.randn(B, C // 2, T, 7, 7, device=dev) * 0.05
gk = torch.randn(B, C // 2, T, 7, 7, device=dev) * 0.05
sup = torch.randn(B, C, T, H, H, device=dev)
u = torch.rand(B, T, T, device=dev)
dy = torch.randn(B, C, T, H, H, device=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(n):
    gq_, gk_, sup_ = gq.clone().requires_grad_(True), gk.clone().requires_grad_(True), sup.clone().requires_grad_(True)
    y = GF.graph_core(gq_, gk_, sup_, u, alpha=0.5, max_hop=3, temperature=1.0)[0]
    y.backward(dy)
torch.cuda.synchronize()
print("ok")
