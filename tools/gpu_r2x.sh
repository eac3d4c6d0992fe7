#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q -k "graph" > gpurun_out/r2_pytest22.log 2>&1; grep -E "^FAILED|passed|failed" gpurun_out/r2_pytest22.log | head; grep -E "^E  " gpurun_out/r2_pytest22.log | head -10
timeout 100 python tools/graph_bwd_ab.py 2>&1 | tail -2
GCA_GRAPH_NOEARLY=1 timeout 100 python tools/graph_bwd_ab.py 2>&1 | tail -1
