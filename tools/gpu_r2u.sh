#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q -k "graph" > gpurun_out/r2_pytest20.log 2>&1; grep -E "^FAILED|passed|failed" gpurun_out/r2_pytest20.log | head; grep -E "^E  " gpurun_out/r2_pytest20.log | head -10
timeout 400 python tools/bench_kernels.py 2>/dev/null | grep "graph" | cut -c1-260
