#!/bin/bash
# One-GPU check run (gpurun -- bash tools/gpu_checks.sh): the full GPU suite, smoke(), a short bench line, the step timelines
# and the stand-alone probes.  The evidence that goes into profiles/ comes from tools/gpu_r02_evidence.sh (1 GPU) and
# tools/gpu_r02_multi.sh N (N GPUs); tools/summarize_profiles.py turns their raw ncu output into the committed summaries.
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; grep -E "^FAILED|passed|failed" gpurun_out/pytest_gpu.log | head; grep -E "^E  " gpurun_out/pytest_gpu.log | head -20
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 2000 --warmup 50 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "ms_per_step_isolated", "gpu_launches")}, "e2e", d["e2e"]["ms_per_step"])
print("roofline", {k: d["roofline"][k] for k in ("frac", "kernel_ms", "step_frac")})
for k, v in (d.get("secondary") or {}).items():
    print(k, {a: b for a, b in v.items() if a != "what"})
PY
timeout 120 python tools/step_timeline.py 2>&1 | tail -12
HOSTIO=1 timeout 120 python tools/step_timeline.py 2>&1 | tail -12
timeout 200 python tools/eager_period.py
timeout 200 python tools/bank_probe.py 2>&1 | tail -1
