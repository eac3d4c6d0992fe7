#!/bin/bash
timeout 300 python -m pytest tests -m gpu -q -x -k "tc32" > gpurun_out/r2_pytest17.log 2>&1; tail -25 gpurun_out/r2_pytest17.log
timeout 120 python - <<'PY'
import sys, torch, torch.nn.functional as F
sys.path.insert(0, "video-graph-ssl_b200")
from gca_b200 import functional as GF
B, K = 256, 65536
mem = F.normalize(torch.randn(K, 128)).cuda(); q = F.normalize(torch.randn(B, 128)).cuda(); k = F.normalize(torch.randn(B, 128)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for algo in ("tc32", "ffma"):
    ts = []
    for i in range(12):
        flush.fill_(i & 1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = GF.infonce_forward(q, k, mem, 0.07, algo=algo, want_grad=True); b.record()
        torch.cuda.synchronize()
        if i >= 4: ts.append(a.elapsed_time(b))
    print(algo, "eager call ms (cold L2):", sum(ts) / len(ts), "loss", float(r["loss"]))
PY
