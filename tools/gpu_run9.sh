cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1000 --warmup 20 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','replicas_consistent','clocks','gpu_launches','sharded_k1m'): print(k, d[k])
    print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'graph', d['config']['cuda_graph'])
except Exception as e: print('no json', e)
PY
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_n2.err | tail -n 8
