#!/bin/bash
for z in 1 0; do
GCA_BENCH_ZC_IN=$z timeout 300 python bench.py --steps 500 --warmup 20 --no-secondary --no-cpu > gpurun_out/r2_bench_zc$z.json 2> gpurun_out/r2_bench_zc$z.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_zc$z.json').read().strip().splitlines()[-1]); print('zc_in=$z', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
done
