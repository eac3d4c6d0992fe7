#!/bin/bash
# flag hand-off prep -> stream (no griddepcontrol.wait in the sweep): full suite + bench + timeline at 1 GPU
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest25.log 2>&1; grep -E "^FAILED|passed|failed" gpurun_out/r2_pytest25.log | head; grep -E "^E  " gpurun_out/r2_pytest25.log | head -20
timeout 300 python bench.py --steps 200 --warmup 20 --no-secondary > gpurun_out/r2_bench_flag.json 2> gpurun_out/r2_bench_flag.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_flag.json').read().strip().splitlines()[-1]); print('flag', d['ms_per_step'], d['ms_per_step_isolated'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['ms_per_step'])"
timeout 120 python tools/step_timeline.py 2>&1 | tail -4
