#!/bin/bash
nvidia-smi topo -m 2>/dev/null | head -6; lscpu | grep -i "numa\|^CPU(s)" | head -6
for nb in 0 1 0 1; do
GCA_BENCH_NO_BIND=$nb timeout 300 python bench.py --steps 500 --warmup 20 --no-secondary --no-cpu > gpurun_out/r2_bench_bind$nb.json 2> gpurun_out/r2_bench_bind$nb.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_bind$nb.json').read().strip().splitlines()[-1]); print('no_bind=$nb', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['cpu_affinity'])"
done
