#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest29.log 2>&1; grep -E "^FAILED|passed|failed" gpurun_out/r2_pytest29.log | head; grep -E "^E  " gpurun_out/r2_pytest29.log | head -20
timeout 300 python bench.py --steps 2000 --warmup 50 --no-secondary > gpurun_out/r2_bench_sp.json 2> gpurun_out/r2_bench_sp.err; tail -3 gpurun_out/r2_bench_sp.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_sp.json').read().strip().splitlines()[-1]); print('selfprep', d['ms_per_step'], d['ms_per_step_isolated'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['step_frac'], 'e2e', d['e2e']['ms_per_step'], d['gpu_launches'])"
timeout 200 python tools/eager_period.py
GCA_X_SELFPREP=0 timeout 200 python tools/eager_period.py
timeout 100 python tools/step_timeline.py 2>&1 | tail -12
