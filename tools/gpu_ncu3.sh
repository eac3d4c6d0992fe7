cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"graph_smem_kernel" -s 12 -c 1 -o gpurun_out/prof_gs \
    python tools/bench_kernels.py > gpurun_out/ncu3.log 2>&1
echo "ncu exit $?"
