cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/bench_kernels.py > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"graph|pairdots|adj_from|negcos|enqueue|row_topk|sim_gemm|row_inv" -c 400 --csv --log-file gpurun_out/launches_secondary.csv \
    python tools/bench_kernels.py > gpurun_out/ncu2.log 2>&1
echo "ncu exit $?"
