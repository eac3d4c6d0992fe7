#!/bin/bash
timeout 120 python tools/step_timeline.py 2>&1 | tail -12
timeout 300 python bench.py --steps 200 --warmup 20 --no-secondary > gpurun_out/r2_bench_flag.json 2> gpurun_out/r2_bench_flag.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_flag.json').read().strip().splitlines()[-1]); print('flag', d['ms_per_step'], d['ms_per_step_isolated'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['ms_per_step'])"
GCA_X_NOFLAG=1 timeout 300 python bench.py --steps 200 --warmup 20 --no-secondary > gpurun_out/r2_bench_noflag.json 2> gpurun_out/r2_bench_noflag.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_noflag.json').read().strip().splitlines()[-1]); print('noflag', d['ms_per_step'], d['ms_per_step_isolated'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['ms_per_step'])"
