#!/bin/bash
# round-2 check: full GPU suite on 2 GPUs, N=2 replica bench, N=1 bench, stream-kernel timeline
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest12.log 2>&1; tail -5 gpurun_out/r2_pytest12.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 300 --warmup 20 --no-sharded > gpurun_out/r2_bench_n2b.json 2> gpurun_out/r2_bench_n2b.err; tail -3 gpurun_out/r2_bench_n2b.err
python bench.py --steps 300 --warmup 20 > gpurun_out/r2_bench_n1b.json 2> gpurun_out/r2_bench_n1b.err; tail -3 gpurun_out/r2_bench_n1b.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_n2b.json", "gpurun_out/r2_bench_n1b.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "ms_per_step_isolated", "replicas_consistent")}, d["e2e"]["ms_per_step"], d.get("roofline", {}).get("kernel_ms"))
    except Exception as e:
        print(f, "ERR", e)
PY
python tools/tc_timeline.py > gpurun_out/r2_tl8.txt 2>&1; cat gpurun_out/r2_tl8.txt
python tools/step_timeline.py > gpurun_out/r2_stl8.txt 2>&1; cat gpurun_out/r2_stl8.txt
