#!/bin/bash
# smallest end-to-end sanity on one GPU: smoke() and the bank tests (argument: extra pytest -k expression)
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python -m pytest tests -m gpu -q -x -k "${1:-bank}" 2>&1 | tail -2
