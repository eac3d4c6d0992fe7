#!/bin/bash
# smallest end-to-end sanity on one GPU: smoke() and one short bench line
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python bench.py --steps 500 --warmup 20 --no-cpu > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err; echo "bench exit $?"; tail -2 gpurun_out/bench_short.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_short.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["ms_per_step"], "pretrain", (d.get("pretrain_clips_per_s") or {}).get("clips_per_s"), list((d.get("secondary") or {}).keys()))
PY
