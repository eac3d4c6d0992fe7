#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q -k "graph or peer" > gpurun_out/r2_pytest21.log 2>&1; grep -E "^FAILED|passed|failed" gpurun_out/r2_pytest21.log | head; grep -E "^E  " gpurun_out/r2_pytest21.log | head -10
timeout 100 python tools/graph_bwd_ab.py 2>&1 | tail -2
