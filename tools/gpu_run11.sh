cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "graph" > gpurun_out/pytest_graph.log 2>&1; echo "exit $?" >> gpurun_out/pytest_graph.log; tail -n 4 gpurun_out/pytest_graph.log
timeout 600 python tools/bench_kernels.py 2>gpurun_out/kernels.err | head -n 8 | tee gpurun_out/kernels_graph.jsonl; tail -n 3 gpurun_out/kernels.err
