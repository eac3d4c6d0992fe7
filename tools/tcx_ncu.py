#!/usr/bin/env python
"""Bring-up aid: a handful of launches of the InfoNCE stream kernel for ncu (not part of the product or the tests)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-graph-ssl_b200"))
import torch
import torch.nn.functional as F
from gca_b200 import _lib, functional as GF

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
K = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
torch.manual_seed(0)
mem = F.normalize(torch.randn(K, 128)).to(torch.bfloat16).cuda()
q, k = F.normalize(torch.randn(B, 128)).cuda(), F.normalize(torch.randn(B, 128)).cuda()
ws = GF.workspace(q.device, GF.infonce_workspace_bytes(B, K, 128, 1, "tcgen05"), "t")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for it in range(12):
    _lib.call("gca_infonce_partials", _lib.ptr(q), _lib.ptr(k), _lib.ptr(mem), 1, B, K, 128, 1 / 0.07, 2, 1, _lib.ptr(ws),
              ws.numel(), st)
torch.cuda.synchronize()
print("ok")
