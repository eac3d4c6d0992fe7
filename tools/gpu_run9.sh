cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 500 --warmup 20 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?"
cat gpurun_out/bench_n2.json | cut -c1-3000; grep -v "^\*\|OMP_NUM" gpurun_out/bench_n2.err | tail -n 12
