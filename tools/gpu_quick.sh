#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; grep -E "^FAILED|passed|failed" gpurun_out/pytest_gpu.log | head; grep -E "^E  " gpurun_out/pytest_gpu.log | head -20
