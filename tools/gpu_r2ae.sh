#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest28.log 2>&1; grep -E "^FAILED|passed|failed" gpurun_out/r2_pytest28.log | head; grep -E "^E  " gpurun_out/r2_pytest28.log | head -20
timeout 300 python bench.py --steps 2000 --warmup 50 --no-secondary > gpurun_out/r2_bench_plan.json 2> gpurun_out/r2_bench_plan.err; tail -3 gpurun_out/r2_bench_plan.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_plan.json').read().strip().splitlines()[-1]); print('plan', d['ms_per_step'], d['ms_per_step_isolated'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['step_frac'], 'e2e', d['e2e']['ms_per_step'], d['wall_s_total'], d['config']['launch_plan'])"
timeout 200 python tools/eager_period.py
