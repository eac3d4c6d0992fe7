#!/usr/bin/env python
"""Back-to-back period of the head step (B=256, K=65536, d=128, bf16 queue) over a pool of 12 queue replicas (as bench.py):
CUDA-graph replays against direct C-ABI calls (3 stream launches per step, programmatic dependent launch between them and
across steps).  One event pair around N steps; also the host-side submit time per step."""
import ctypes, os, sys, time, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-graph-ssl_b200"))
import gca_b200
from gca_b200.graphed import GraphedMoCoStep

B, K, D, POOL = 256, 65536, 128, 12
dev = torch.device("cuda", 0)
torch.manual_seed(1)
steps = []
for i in range(POOL):
    moco = gca_b200.RGBMoCo(D, K=K, T=0.07, queue_dtype="bf16").to(dev)
    s = GraphedMoCoStep(moco, B, B, want_rank=False)
    s.inputs.copy_(torch.nn.functional.normalize(torch.randn(3 * B, D, device=dev)))
    s.capture()
    steps.append(s)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

def run(fn, n):
    for i in range(100):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for i in range(n):
        fn(i)
    b.record()
    t_submit = time.perf_counter() - t0
    torch.cuda.synchronize()
    return {"period_us": round(a.elapsed_time(b) / n * 1e3, 2), "submit_us_per_step": round(t_submit / n * 1e6, 2)}

out = {}
for rep in range(2):
    out["graph_%d" % rep] = run(lambda i: steps[i % POOL].graph.replay(), 3000)
    out["eager_%d" % rep] = run(lambda i: steps[i % POOL]._enqueue_work(st), 3000)
    out["plan_%d" % rep] = run(lambda i: steps[i % POOL].step(), 3000)
out["loss"] = [float(steps[0].loss), float(steps[1].loss)]
print(json.dumps(out))
