#!/bin/bash
# Round-2 evidence run (1 GPU): bench line, ncu launch list of the same command, one ncu --set full capture of the head kernels
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench exit $?"
timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu --no-secondary > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 30 --warmup 3 --no-cpu --no-secondary > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu --no-secondary > gpurun_out/plain_b.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"infonce_tcx|infonce_finalize|infonce_prep" -s 60 -c 6 -o gpurun_out/r02_prof_head \
    python bench.py --steps 30 --warmup 3 --no-cpu --no-secondary > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out/r02_* 2>/dev/null
