cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/bench_kernels.py > gpurun_out/kernels.jsonl 2> gpurun_out/kernels.err; echo "kernels exit $?"
cat gpurun_out/kernels.jsonl; tail -n 5 gpurun_out/kernels.err
