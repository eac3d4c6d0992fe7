#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest14.log 2>&1; tail -4 gpurun_out/r2_pytest14.log
timeout 120 python tools/tc_timeline.py > gpurun_out/r2_tl8.txt 2>&1; cat gpurun_out/r2_tl8.txt
timeout 120 python tools/step_timeline.py > gpurun_out/r2_stl8.txt 2>&1; cat gpurun_out/r2_stl8.txt
