#!/bin/bash
# 2 GPUs: key push inside the prep launch with late trigger (linear launch chain) against the side-stream push
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=2
timeout 600 python -m pytest tests -m gpu -q -x -k "peer or replica or exchange or nccl" > gpurun_out/r2_pytest31.log 2>&1; grep -E "^FAILED|passed|failed" gpurun_out/r2_pytest31.log | head; grep -E "^E  " gpurun_out/r2_pytest31.log | head -20
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 tests/peer_exchange_worker.py > gpurun_out/r2_peer_exchange_worker_n$N.log 2>&1
echo "peer_exchange_worker exit $?"; grep -E "_OK|Error" gpurun_out/r2_peer_exchange_worker_n$N.log | head -3
for mode in late side lategraph; do
  unset GCA_PUSH_SIDE GCA_BENCH_PREFER_GRAPH
  if [ $mode = side ]; then export GCA_PUSH_SIDE=1; fi
  if [ $mode = lategraph ]; then export GCA_BENCH_PREFER_GRAPH=1; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 1000 --warmup 20 --no-sharded > gpurun_out/r2_bench_n2_$mode.json 2> gpurun_out/r2_bench_n2_$mode.err
  echo "bench $mode exit $?"; grep -v "OMP_NUM\|^\*\*\*" gpurun_out/r2_bench_n2_$mode.err | tail -2
  python - $mode <<'PY'
import json, sys
d = json.loads(open("gpurun_out/r2_bench_n2_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], {k: d.get(k) for k in ("value", "ms_per_step", "ms_per_step_isolated", "replicas_consistent", "gpu_launches")}, "e2e", d["e2e"]["ms_per_step"], d["config"].get("launch_plan"))
PY
done
