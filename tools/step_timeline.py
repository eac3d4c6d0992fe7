#!/usr/bin/env python
"""Bring-up aid: globaltimer timeline of one graph-replayed head step (prep -> stream -> finalize)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-graph-ssl_b200"))
import torch
import torch.nn.functional as F
buf = torch.zeros(32 * 1024, dtype=torch.int64, device="cuda")
os.environ["GCA_TC_TIMEBUF"] = hex(buf.data_ptr())
import gca_b200
from gca_b200.graphed import GraphedMoCoStep
B, K = 256, 65536
moco = gca_b200.RGBMoCo(128, K=K, queue_dtype="bf16").cuda()
step = GraphedMoCoStep(moco, B, B).capture()
HOSTIO = os.environ.get("HOSTIO") == "1"              # the e2e variant: q | k | all_k read from / results stored to pinned host memory
if HOSTIO:
    host_in = torch.nn.functional.normalize(torch.randn(3 * B, 128)).pin_memory()
    host_out = torch.empty_like(step.outputs, device="cpu").pin_memory()
    step.capture_host_io(host_in, host_out, zero_copy_out=True, zero_copy_in=True)
    step.graph = step.graph_io
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
BIG = 1 << 62
for cold in (True, False):
    for it in range(200):
        if cold:
            flush.fill_(it & 1)
        step.graph.replay()
    if cold:
        flush.fill_(1)
    buf.zero_(); buf[32 * 1000 + 0] = BIG; buf[32 * 1000 + 2] = BIG
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); step.graph.replay(); b.record()
    torch.cuda.synchronize()
    t = buf.cpu()
    cta = t[:32 * 148].view(148, 32).double()
    s_in, s_out = cta[:, 0].min(), cta[:, 8].max()
    f_in, f_out, p_in = float(t[32 * 1000]), float(t[32 * 1000 + 1]), float(t[32 * 1000 + 2])
    print("%s L2: event time %.2f us | prep entry 0.00 | stream first entry %.2f, last exit %.2f (span %.2f) | finalize first entry %.2f, "
          "last-block exit %.2f (span %.2f)" % ("cold" if cold else "warm", a.elapsed_time(b) * 1e3, (s_in - p_in) / 1e3, (s_out - p_in) / 1e3,
          (s_out - s_in) / 1e3, (f_in - p_in) / 1e3, (f_out - p_in) / 1e3, (f_out - f_in) / 1e3))

    for sl, nm in ((0, "stream CTA entry"), (9, "producer: first TMA issued"), (10, "mma: q block seen"), (11, "mma: first tile landed"),
                   (4, "first S tile seen"), (5, "main loop done"), (6, "last O GEMM done"), (7, "partials written"), (8, "exit")):
        rel = (cta[:, sl] - p_in) / 1e3
        print("   stream %-28s mean %6.2f  min %6.2f  max %6.2f us since prep entry" % (nm, rel.mean(), rel.min(), rel.max()))
    print("   prep last row CTA done %.2f us" % ((float(t[32 * 1000 + 4]) - p_in) / 1e3))
    if float(cta[:, 14].max()) > 0:      # fused finalize stamps (per CTA): 12 partial written, 13 past the grid barrier, 14 done
        for sl, nm in ((5, "main loop done"), (6, "last MMA done"), (12, "partial written"), (13, "past grid barrier"), (14, "fused finalize done"), (8, "exit")):
            rel = (cta[:, sl] - p_in) / 1e3
            print("   %-22s mean %6.2f  min %6.2f  max %6.2f us since prep entry" % (nm, rel.mean(), rel.min(), rel.max()))
        continue
    fb = t[32 * 400: 32 * 400 + 4 * 256].view(256, 4).double()
    e = (fb[:, 0] - p_in) / 1e3; st_ = (fb[:, 1] - p_in) / 1e3; ac = (fb[:, 2] - p_in) / 1e3
    print("   finalize row blocks: entry min %.2f max %.2f | stats+loads done min %.2f max %.2f | accumulate done min %.2f max %.2f | "
          "last block after ticket %.2f" % (e.min(), e.max(), st_.min(), st_.max(), ac.min(), ac.max(), (float(t[32 * 1000 + 6]) - p_in) / 1e3))
