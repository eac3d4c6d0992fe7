#!/usr/bin/env python
"""Runs the config-5 retrieval (3783 x 13320 x 512, k=50) a few times (for an ncu launch list / per-kernel times)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-graph-ssl_b200"))
from gca_b200 import functional as GF
torch.manual_seed(0)
qry = torch.randn(3783, 512, device="cuda")
gal = torch.randn(13320, 512, device="cuda")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    idx, val = GF.cosine_topk(qry, gal, 50)
torch.cuda.synchronize()
print("ok", int(idx[0, 0]))
