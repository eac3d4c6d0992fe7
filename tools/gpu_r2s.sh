#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest18.log 2>&1; grep -E "^FAILED|passed|failed" gpurun_out/r2_pytest18.log | head -30
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
