cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "graph" > gpurun_out/pytest_graph.log 2>&1; echo "exit $?" >> gpurun_out/pytest_graph.log; tail -n 4 gpurun_out/pytest_graph.log
timeout 600 python tools/bench_kernels.py > gpurun_out/kernels.jsonl 2> gpurun_out/kernels.err; echo "kernels exit $?"
cat gpurun_out/kernels.jsonl; tail -n 5 gpurun_out/kernels.err
