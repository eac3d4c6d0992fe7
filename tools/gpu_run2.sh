set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log
tail -n 15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?" >> gpurun_out/smoke.log
tail -n 5 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
cat gpurun_out/bench_n1.json; tail -n 20 gpurun_out/bench_n1.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:infonce_tc -s 10 -c 3 -o gpurun_out/prof_tc \
    python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
tail -n 5 gpurun_out/ncu_list.log gpurun_out/ncu_full.log 2>/dev/null | tail -n 30
ls -la gpurun_out
