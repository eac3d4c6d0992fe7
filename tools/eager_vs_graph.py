#!/usr/bin/env python
"""Device time of one head step (B=256, K=65536, d=128, bf16 queue) launched three ways, L2 flushed before every step:
CUDA-graph replay, direct C-ABI call (3 stream launches), and the same with the launches issued from a C loop-free
python call but WITHOUT the flush slack (back to back) to see whether the CPU keeps up."""
import ctypes, os, sys, time, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-graph-ssl_b200"))
import gca_b200
from gca_b200.graphed import GraphedMoCoStep

B, K, D = 256, 65536, 128
dev = torch.device("cuda", 0)
torch.manual_seed(1)
moco = gca_b200.RGBMoCo(D, K=K, T=0.07, queue_dtype="bf16").to(dev)
s = GraphedMoCoStep(moco, B, B)
s.inputs.copy_(torch.nn.functional.normalize(torch.randn(3 * B, D, device=dev)))
s.capture()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

def run(fn, n, do_flush):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for i in range(50):
        if do_flush: flush.fill_(i & 1)
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        if do_flush: flush.fill_(i & 1)
        ev[i][0].record(); fn(); ev[i][1].record()
    t_submit = time.perf_counter() - t0
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    return {"mean_us": round(sum(ms) / n * 1e3, 2), "median_us": round(ms[n // 2] * 1e3, 2), "submit_us_per_step": round(t_submit / n * 1e6, 2),
            "wall_us_per_step": round(wall / n * 1e6, 2)}

out = {}
out["graph_flush"] = run(lambda: s.graph.replay(), 2000, True)
out["eager_flush"] = run(lambda: s._enqueue_work(st), 2000, True)
out["graph_b2b"] = run(lambda: s.graph.replay(), 2000, False)
out["eager_b2b"] = run(lambda: s._enqueue_work(st), 2000, False)
print(json.dumps(out))
