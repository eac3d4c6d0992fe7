# Round-end evidence run, part A (1 GPU): tests, smoke, bench, reference arm, secondary kernels, timelines, ncu launch list
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log; tail -n 3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?" >> gpurun_out/smoke.log; tail -n 2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 60 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_n1.err; echo "ref exit $?"
timeout 400 python tools/bench_kernels.py > gpurun_out/kernels.jsonl 2>> gpurun_out/bench_n1.err; echo "kernels exit $?"
timeout 300 python tools/pretrain_step.py --steps 20 --warmup 5 2>/dev/null | grep workload > gpurun_out/pretrain_n1.jsonl; echo "pretrain exit $?"
timeout 120 python tools/step_timeline.py > gpurun_out/step_timeline.txt 2>&1
timeout 120 python tools/tc_timeline.py 256 65536 > gpurun_out/tc_timeline.txt 2>&1
timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 30 --warmup 3 --no-cpu > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
ls -la gpurun_out | head -30
