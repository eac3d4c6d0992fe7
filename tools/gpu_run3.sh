set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 120 python tools/tc_bringup.py 256 65536 > gpurun_out/tc3.log 2>&1; echo "tc3 exit $?" >> gpurun_out/tc3.log; cat gpurun_out/tc3.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log
tail -n 25 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?" >> gpurun_out/smoke.log
tail -n 5 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
cat gpurun_out/bench_n1.json; tail -n 20 gpurun_out/bench_n1.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"infonce_tc|finalize" -s 20 -c 4 -o gpurun_out/prof_tc \
    python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
