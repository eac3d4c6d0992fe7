#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q -k "graph" > gpurun_out/r2_pytest19.log 2>&1; grep -E "^FAILED|passed|failed" gpurun_out/r2_pytest19.log | head -30; grep -E "^E  " gpurun_out/r2_pytest19.log | head -20
