cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"infonce_tc" -s 40 -c 2 -o gpurun_out/prof_tc \
    python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -n 3 gpurun_out/ncu_full.log | cut -c1-300
