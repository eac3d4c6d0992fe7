#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
N=2
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_n$N.log 2>&1; grep -E "^FAILED|passed|failed" gpurun_out/pytest_gpu_n$N.log | head; grep -E "^E  " gpurun_out/pytest_gpu_n$N.log | head -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 1000 --warmup 20 --no-sharded > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench exit $?"; grep -v "OMP_NUM\|^\*\*\*" gpurun_out/bench_n$N.err | tail -2
python - $N <<'PY'
import json, sys
d = json.loads(open("gpurun_out/bench_n%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "ms_per_step_isolated", "replicas_consistent", "gpu_launches")}, "e2e", d["e2e"]["ms_per_step"])
PY
