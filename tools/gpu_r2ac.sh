#!/bin/bash
# 2 GPUs: flag hand-off (key push inside the prep launch, linear graph) against the side-stream push
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=2
timeout 600 python -m pytest tests -m gpu -q -x -k "peer or replica or shard or exchange or shuffle or nccl or graphed" > gpurun_out/r2_pytest26.log 2>&1; grep -E "^FAILED|passed|failed" gpurun_out/r2_pytest26.log | head; grep -E "^E  " gpurun_out/r2_pytest26.log | head -20
for w in sharded_graph_worker peer_exchange_worker; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 tests/$w.py > gpurun_out/r2_${w}_n$N.log 2>&1
  echo "$w exit $?"; grep -E "_OK|Error" gpurun_out/r2_${w}_n$N.log | head -3
done
for mode in flag noflag; do
  if [ $mode = noflag ]; then export GCA_X_NOFLAG=1; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 300 --warmup 20 --no-sharded > gpurun_out/r2_bench_n2_$mode.json 2> gpurun_out/r2_bench_n2_$mode.err
  echo "bench $mode exit $?"; tail -2 gpurun_out/r2_bench_n2_$mode.err
  python - $mode <<'PY'
import json, sys
d = json.loads(open("gpurun_out/r2_bench_n2_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], {k: d.get(k) for k in ("value", "ms_per_step", "ms_per_step_isolated", "replicas_consistent", "gpu_launches")}, "e2e", d["e2e"]["ms_per_step"])
PY
done
