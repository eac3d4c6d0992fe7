#!/usr/bin/env python
"""Bring-up aid: time the InfoNCE stream kernel (train of launches over distinct queue copies, cold L2) and the captured
head step for the library named by GCA_B200_LIB (default: the in-tree build).  Not part of the product or the tests."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-graph-ssl_b200"))
import torch
import torch.nn.functional as F
import gca_b200
from gca_b200 import _lib, functional as GF
from gca_b200.graphed import GraphedMoCoStep

B = int(os.environ.get("TCX_B", 256)); K = int(os.environ.get("TCX_K", 65536)); D = 128; T = 0.07
dev = torch.device("cuda", 0)
torch.manual_seed(1)
moco = gca_b200.RGBMoCo(D, K=K, T=T, queue_dtype="bf16").to(dev)
q, k = F.normalize(torch.randn(B, D)).to(dev), F.normalize(torch.randn(B, D)).to(dev)
ws = GF.workspace(dev, GF.infonce_workspace_bytes(B, K, D, 1, "tcgen05"), "bench")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
_lib.call("gca_infonce_partials", _lib.ptr(q), _lib.ptr(k), _lib.ptr(moco.memory), 1, B, K, D, 1.0 / T, 2, 1, _lib.ptr(ws), ws.numel(), st())
torch.cuda.synchronize()
NQ = max(2, min(12, (220 << 20) // (K * D * 2)))
queues = [moco.memory.clone() for _ in range(NQ)]
tg = torch.cuda.CUDAGraph()
ta, tb = torch.cuda.Event(enable_timing=True, external=True), torch.cuda.Event(enable_timing=True, external=True)
with torch.cuda.graph(tg):
    ta.record()
    for qu in queues:
        _lib.call("gca_infonce_partials", _lib.ptr(q), _lib.ptr(k), _lib.ptr(qu), 1, B, K, D, 1.0 / T, 2, 3, _lib.ptr(ws), ws.numel(), st())
    tb.record()
tt = []
for i in range(35):
    flush.fill_(i & 1)
    tg.replay()
    torch.cuda.synchronize()
    if i >= 5:
        tt.append(ta.elapsed_time(tb) / NQ)
k_us = 1e3 * sum(tt) / len(tt)
# whole step (prep + stream + finalize + enqueue) as one captured graph, cold L2, per-step events
state = torch.tensor([0, 0], dtype=torch.int64, device=dev)
steps = []
for i in range(4):
    s = GraphedMoCoStep(moco, B, B, state=state, want_rank=os.environ.get('TCX_NORANK') != '1')
    s.inputs.copy_(torch.cat([F.normalize(torch.randn(B, D)), F.normalize(torch.randn(B, D)), F.normalize(torch.randn(B, D))]).to(dev).view_as(s.inputs))
    s.capture()
    steps.append(s)
ts = []
for i in range(60):
    flush.fill_(i & 1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); steps[i % 4].step(); b.record()
    torch.cuda.synchronize()
    if i >= 10:
        ts.append(a.elapsed_time(b))
ts.sort()
print(json.dumps({"lib": os.environ.get("GCA_B200_LIB", "default"), "B": B, "K": K, "kernel_us_train": round(k_us, 2),
                  "tflops": round(4.0 * B * K * D / k_us / 1e6, 1), "step_us_mean": round(1e3 * sum(ts) / len(ts), 2),
                  "step_us_median": round(1e3 * ts[len(ts) // 2], 2), "loss": float(steps[0].loss)}))

# back-to-back: NQ captured steps, each over its OWN queue copy (NQ x 16.8 MB > L2), replayed round-robin between ONE event pair
NQ2 = 12
mocos = [gca_b200.RGBMoCo(D, K=K, T=T, queue_dtype="bf16").to(dev) for _ in range(NQ2)]
gs = []
for m in mocos:
    s = GraphedMoCoStep(m, B, B, want_rank=os.environ.get('TCX_NORANK') != '1')
    s.inputs.copy_(torch.cat([F.normalize(torch.randn(B, D)), F.normalize(torch.randn(B, D)), F.normalize(torch.randn(B, D))]).to(dev).view_as(s.inputs))
    s.capture()
    gs.append(s)
for rep in range(3):
    flush.fill_(rep)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for i in range(240):
        gs[i % NQ2].graph.replay()
    b.record()
    torch.cuda.synchronize()
    bb = a.elapsed_time(b) / 240 * 1e3
print(json.dumps({"back_to_back_step_us": round(bb, 2), "loss": float(gs[0].loss)}))
