# A/B of the key exchange at N GPUs: fused peer-memory exchange vs NCCL all-gather (bench.py, replicas only)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/peer_exchange_worker.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -n 25
echo "worker exit ${PIPESTATUS[0]}"
for mode in fused nccl; do
GCA_BENCH_EXCHANGE=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 1000 --warmup 20 --no-sharded > gpurun_out/bench_n${N}_$mode.json 2> gpurun_out/bench_n${N}_$mode.err; echo "bench n$N $mode exit $?"
python - $N $mode <<'PY'
import json, sys
n, mode = sys.argv[1], sys.argv[2]
try:
    d=json.loads(open('gpurun_out/bench_n%s_%s.json'%(n,mode)).read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','replicas_consistent','gpu_launches'): print(k, d[k])
    print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'graph', d['config']['cuda_graph'], d['config']['key_exchange'])
except Exception as e: print('no json', e)
PY
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_n${N}_$mode.err | tail -n 8
done
