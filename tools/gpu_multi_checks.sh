#!/bin/bash
# N-GPU check run (gpurun --gpus N -- bash tools/gpu_multi_checks.sh N): the GPU suite (its multi-GPU tests run when >= 2 GPUs are
# visible), the torchrun workers and a short bench line
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_n$N.log 2>&1; grep -E "^FAILED|passed|failed" gpurun_out/pytest_gpu_n$N.log | head; grep -E "^E  " gpurun_out/pytest_gpu_n$N.log | head -20
for w in sharded_graph_worker peer_exchange_worker; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 tests/$w.py > gpurun_out/${w}_n$N.log 2>&1
  echo "$w exit $?"; grep -E "_OK|Error" gpurun_out/${w}_n$N.log | head -3
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 1000 --warmup 20 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench exit $?"; grep -v "OMP_NUM\|^\*\*\*" gpurun_out/bench_n$N.err | tail -2
python - $N <<'PY'
import json, sys
d = json.loads(open("gpurun_out/bench_n%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "ms_per_step_isolated", "replicas_consistent", "gpu_launches")}, "e2e", d["e2e"]["ms_per_step"])
for k in ("sharded_k1m", "sharded_k1m_strong"):
    s = d.get(k)
    if s: print(k, s["ms_per_step"], s.get("matches_eager_path"), s.get("nccl_variant"))
print("pretrain", (d.get("pretrain_clips_per_s") or {}).get("clips_per_s"))
PY
