#!/bin/bash
# cluster-pair merge of the split partials: full GPU suite, A/B bench, step timeline
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest24.log 2>&1; grep -E "^FAILED|passed|failed" gpurun_out/r2_pytest24.log | head; grep -E "^E  " gpurun_out/r2_pytest24.log | head -20
timeout 300 python bench.py --steps 200 --warmup 20 --no-secondary > gpurun_out/r2_bench_pair.json 2> gpurun_out/r2_bench_pair.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_pair.json').read().strip().splitlines()[-1]); print('pair', d['ms_per_step'], d['roofline'], d['e2e'])"
GCA_X_NOPAIR=1 timeout 300 python bench.py --steps 200 --warmup 20 --no-secondary > gpurun_out/r2_bench_nopair.json 2> gpurun_out/r2_bench_nopair.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_nopair.json').read().strip().splitlines()[-1]); print('nopair', d['ms_per_step'], d['roofline'], d['e2e'])"
timeout 120 python tools/step_timeline.py 2>&1 | tail -4
GCA_X_NOPAIR=1 timeout 120 python tools/step_timeline.py 2>&1 | tail -4
