#!/usr/bin/env python
"""Bring-up aid: where does a head step's time go?  CUDA-event timings (mean over 50, cold = L2 flushed before each) of the
pieces of the step.  Not part of the product or the tests."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-graph-ssl_b200"))
import torch
import torch.nn.functional as F
import gca_b200
from gca_b200 import _lib, functional as GF
from gca_b200.graphed import GraphedMoCoStep

B, K = 256, 65536
torch.manual_seed(0)
moco = gca_b200.RGBMoCo(128, K=K, queue_dtype="bf16").cuda()
q, k = F.normalize(torch.randn(B, 128)).cuda(), F.normalize(torch.randn(B, 128)).cuda()
ws = GF.workspace(q.device, GF.infonce_workspace_bytes(B, K, 128, 1, "tcgen05"), "t")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
step = GraphedMoCoStep(moco, B, B).capture()
step.q.copy_(q); step.k.copy_(k); step.all_k.copy_(k)
out = {n: torch.empty(B, device="cuda") for n in ("lr", "lse", "pos")}
rank = torch.empty(B, dtype=torch.int32, device="cuda"); hits = torch.empty(2, dtype=torch.int32, device="cuda")
loss = torch.empty((), device="cuda"); dq = torch.empty(B, 128, device="cuda")


def partials(flag):
    _lib.call("gca_infonce_partials", _lib.ptr(q), _lib.ptr(k), _lib.ptr(moco.memory), 1, B, K, 128, 1 / 0.07, 2, flag, _lib.ptr(ws), ws.numel(), st)


def fwd():
    _lib.call("gca_infonce_fwd", _lib.ptr(q), _lib.ptr(k), _lib.ptr(moco.memory), 1, B, K, 128, 1 / 0.07, 2, _lib.ptr(loss), _lib.ptr(out["lr"]),
              _lib.ptr(out["lse"]), _lib.ptr(out["pos"]), _lib.ptr(rank), _lib.ptr(hits), _lib.ptr(dq), None, _lib.ptr(ws), ws.numel(), st)


cases = {"stream kernel only (eager)": lambda: partials(3), "prep + stream (eager)": lambda: partials(1),
         "gca_infonce_fwd: prep + stream + finalize (eager)": fwd, "graph replay: gca_moco_step": lambda: step.graph.replay(),
         "empty (event pair only)": lambda: None}
for name, fn in cases.items():
    for cold in (True, False):
        ts = []
        for i in range(60):
            if cold:
                flush.fill_(i & 1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            if i >= 10:
                ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        print("%-52s %s L2: mean %6.2f us  median %6.2f  min %6.2f" % (name, "cold" if cold else "warm", sum(ts) / len(ts), ts[len(ts) // 2], ts[0]))
# back-to-back replays without host sync in between (what a training loop does)
for cold in (True, False):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 200
    torch.cuda.synchronize()
    a.record()
    for i in range(n):
        step.graph.replay()
    b.record(); torch.cuda.synchronize()
    print("graph replay x%d back-to-back (warm L2): %.2f us per step" % (n, a.elapsed_time(b) * 1e3 / n))
    break
