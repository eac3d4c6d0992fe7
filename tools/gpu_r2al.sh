#!/bin/bash
for sy in 0 1 0 1; do
GCA_BENCH_E2E_SYNC=$sy timeout 300 python bench.py --steps 1000 --warmup 50 --no-secondary --no-cpu > gpurun_out/r2_bench_done$sy.json 2> gpurun_out/r2_bench_done$sy.err; tail -2 gpurun_out/r2_bench_done$sy.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_done$sy.json').read().strip().splitlines()[-1]); print('e2e_sync=$sy', d['ms_per_step'], d['ms_per_step_isolated'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['loss_last'])"
done
timeout 600 python -m pytest tests -m gpu -q -x -k "host_io or graphed or headline" > gpurun_out/r2_pytest33.log 2>&1; grep -E "^FAILED|passed|failed" gpurun_out/r2_pytest33.log | head -3
