#!/bin/bash
# Round-2 multi-GPU evidence: worker tests (oracle parity of the sharded / replica steps) and the bench line at N GPUs
N=${1:-8}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for w in sharded_graph_worker peer_exchange_worker; do      # (shuffle_bn_worker checks a world-size-2 fixture: pytest runs it at 2 GPUs)
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 tests/$w.py > gpurun_out/r02_${w}_n$N.log 2>&1
  echo "$w exit $?"; grep -E "_OK|Error" gpurun_out/r02_${w}_n$N.log | head -3
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 300 --warmup 20 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
echo "bench exit $?"; tail -2 gpurun_out/r02_bench_n$N.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_bench_n$N.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "ms_per_step_isolated", "replicas_consistent")}, "e2e", d["e2e"]["ms_per_step"])
for k in ("sharded_k1m", "sharded_k1m_strong"):
    s = d.get(k)
    if s: print(k, s["ms_per_step"], s.get("matches_eager_path"), s.get("nccl_variant"))
print("pretrain", d.get("pretrain_clips_per_s"))
PY
