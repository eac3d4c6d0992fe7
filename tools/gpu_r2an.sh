#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q -x -k "bank" > gpurun_out/r2_pytest34.log 2>&1; grep -E "^FAILED|passed|failed" gpurun_out/r2_pytest34.log | head; grep -E "^E  " gpurun_out/r2_pytest34.log | head -20
timeout 200 python tools/bank_probe.py 2>&1 | tail -3
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:bank_ -c 12 python tools/bank_probe.py 2>&1 | grep -E "bank_.*\(|gpu__time|dram__bytes" | head -24
