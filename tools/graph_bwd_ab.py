#!/usr/bin/env python
"""Bring-up aid: time gca_graph_fwd / gca_graph_bwd at BASELINE config 3 (128 x [192, 8, 14, 14], sub_sample) -- cold L2."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-graph-ssl_b200"))
import torch
from gca_b200 import _lib
P = _lib.ptr
Bv, C, T, H, W = 128, 192, 8, 14, 14
Cq, S, HW = C // 2, (H // 2) * (W // 2), H * W
g = torch.Generator(device="cuda").manual_seed(0)
gq = torch.randn(Bv, Cq, T, S, device="cuda", generator=g) * 0.1; gk = torch.randn(Bv, Cq, T, S, device="cuda", generator=g) * 0.1
sup = torch.randn(Bv, C, T, HW, device="cuda", generator=g); u = torch.rand(Bv, T, T, device="cuda", generator=g)
dy = torch.randn(Bv, C, T, HW, device="cuda", generator=g)
sim = torch.empty(Bv, T, T, device="cuda"); adj = torch.empty_like(sim); s_ = torch.empty_like(sim); y = torch.empty_like(sup)
dq, dk, ds = torch.empty_like(gq), torch.empty_like(gk), torch.empty_like(sup)
ws = torch.zeros(int(_lib.load().gca_graph_workspace_bytes(Bv, T)), dtype=torch.uint8, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def fwd(): _lib.call("gca_graph_fwd", P(gq), P(gk), Cq, S, P(sup), C, HW, T, Bv, P(u), 0.5, 3, 1.0, 0, P(sim), P(adj), P(s_), P(y), P(ws), ws.numel(), st)
def bwd(): _lib.call("gca_graph_bwd", P(gq), P(gk), Cq, S, P(sup), C, HW, T, Bv, P(sim), P(adj), P(s_), P(dy), 0.5, 3, 1.0, 0, P(dq), P(dk), P(ds), P(ws), ws.numel(), st)
for name, fn in (("fwd", fwd), ("bwd", bwd)):
    ts = []
    for i in range(25):
        flush.fill_(i & 1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i >= 5: ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    print("%s NOSLAB=%s NOREV=%s: median %.1f us min %.1f" % (name, os.environ.get("GCA_GRAPH_NOSLAB", "0"), os.environ.get("GCA_GRAPH_NOREV", "0"), ts[len(ts) // 2], ts[0]))
