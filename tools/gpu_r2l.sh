#!/bin/bash
# 2-GPU check of the peer-memory K-sharded step
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tests/sharded_graph_worker.py > gpurun_out/r2_shard_w2.log 2>&1; tail -12 gpurun_out/r2_shard_w2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 300 --warmup 20 > gpurun_out/r2_bench_n2c.json 2> gpurun_out/r2_bench_n2c.err; tail -3 gpurun_out/r2_bench_n2c.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_n2c.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step")}, d["e2e"]["ms_per_step"])
print(json.dumps(d.get("sharded_k1m"), indent=1)); print(json.dumps(d.get("sharded_k1m_strong"), indent=1))
PY
