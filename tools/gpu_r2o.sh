#!/bin/bash
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tests/sharded_graph_worker.py > gpurun_out/r2_shard_w2.log 2>&1; tail -4 gpurun_out/r2_shard_w2.log
for bl in 256 128; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29563 tools/shard_timeline.py $bl 2>&1 | grep "rank "
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 300 --warmup 20 > gpurun_out/r2_bench_n2d.json 2> gpurun_out/r2_bench_n2d.err; tail -3 gpurun_out/r2_bench_n2d.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_n2d.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step")}, d["e2e"]["ms_per_step"])
for k in ("sharded_k1m", "sharded_k1m_strong"):
    s = d[k]; print(k, s["ms_per_step"], s["matches_eager_path"], s["nccl_variant"])
PY
