#!/bin/bash
timeout 200 python tools/eager_period.py
GCA_PREP_PDL=0 timeout 200 python tools/eager_period.py
timeout 600 python -m pytest tests -m gpu -q -x -k "infonce or step or moco or headline" > gpurun_out/r2_pytest27.log 2>&1; grep -E "^FAILED|passed|failed" gpurun_out/r2_pytest27.log | head; grep -E "^E  " gpurun_out/r2_pytest27.log | head -20
