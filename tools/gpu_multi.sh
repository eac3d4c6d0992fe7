# Multi-GPU evidence run (gpurun --gpus N -- bash tools/gpu_multi.sh N): parity workers, bench (replicas + K-sharded), pre-train step
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tests/sharded_graph_worker.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -n 12
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/peer_exchange_worker.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -n 25
echo "worker exit ${PIPESTATUS[0]}"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 1000 --warmup 20 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N exit $?"
python - $N <<'PY'
import json, sys
n=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/bench_n%s.json'%n).read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','replicas_consistent','clocks','gpu_launches','sharded_k1m'): print(k, d[k])
    print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'graph', d['config']['cuda_graph'])
except Exception as e: print('no json', e)
PY
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_n$N.err | tail -n 8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tools/pretrain_step.py --steps 10 --warmup 3 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -n 6 | tee gpurun_out/pretrain_n$N.jsonl
