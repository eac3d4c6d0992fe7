#!/bin/bash
# 2-GPU check of the peer-fused replica step (push warp in the stream kernel)
timeout 300 python -m pytest tests -m gpu -q -x -k "multi_gpu or peer or graphed" > gpurun_out/r2_pytest13.log 2>&1; tail -5 gpurun_out/r2_pytest13.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 300 --warmup 20 --no-sharded > gpurun_out/r2_bench_n2b.json 2> gpurun_out/r2_bench_n2b.err; tail -3 gpurun_out/r2_bench_n2b.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_n2b.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "ms_per_step_isolated", "replicas_consistent")}, d["e2e"]["ms_per_step"])
    except Exception as e:
        print(f, "ERR", e)
PY
