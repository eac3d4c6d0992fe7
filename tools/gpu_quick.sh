#!/bin/bash
# one full bench line on one GPU (the committed profiles/r02_bench_n1.json)
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench exit $?"; tail -2 gpurun_out/r02_bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_n1.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "ms_per_step_isolated", "gpu_launches", "steps")}, "e2e", d["e2e"]["ms_per_step"], "cpu", d["cpu_baseline"]["value"])
print("roofline", {k: d["roofline"][k] for k in ("frac", "kernel_ms", "step_frac")})
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 | tail -1 | cut -c1-300
