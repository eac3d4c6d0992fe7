#!/bin/bash
# retrieval: panel-free top-k tests + A/B timing against the panel path
timeout 600 python -m pytest tests -m gpu -q -k "retrieval" > gpurun_out/r2_pytest23.log 2>&1; grep -E "^FAILED|passed|failed" gpurun_out/r2_pytest23.log | head; grep -E "^E  " gpurun_out/r2_pytest23.log | head -20
cat > /tmp/rt.py <<'PY'
import os, sys, torch
sys.path.insert(0, "video-graph-ssl_b200")
from gca_b200 import functional as GF
torch.manual_seed(0)
qry = torch.randn(3783, 512, device="cuda"); gal = torch.randn(13320, 512, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3): GF.cosine_topk(qry, gal, 50)
ts = []
for _ in range(10):
    flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); GF.cosine_topk(qry, gal, 50); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
ts.sort(); print("retrieval c5 ms median %.4f min %.4f  panel=%s" % (ts[len(ts)//2], ts[0], os.environ.get("GCA_SIM_PANEL", "0")))
PY
timeout 120 python /tmp/rt.py
GCA_SIM_PANEL=1 timeout 120 python /tmp/rt.py
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv --log-file gpurun_out/retr_fused_launches.csv python tools/retrieval_probe.py 2 > gpurun_out/retr_ncu.log 2>&1; tail -2 gpurun_out/retr_ncu.log
