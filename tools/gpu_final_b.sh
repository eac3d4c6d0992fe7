# Round-end evidence run, part B (1 GPU): one ncu --set full capture of the head kernels (after the same command ran clean)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu > gpurun_out/plain_b.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"infonce_tc|infonce_finalize|infonce_prep" -s 60 -c 6 -o gpurun_out/prof_r01 \
    python bench.py --steps 30 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out/prof_r01* 2>/dev/null
