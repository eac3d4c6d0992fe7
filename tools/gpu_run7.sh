cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 120 python tools/step_timeline.py 2>&1 | tail -n 4
GCA_NO_PDL=1 timeout 120 python tools/step_timeline.py 2>&1 | tail -n 2
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --no-cpu > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print('PDL   step us', d['ms_per_step']*1e3, 'value', d['value'], 'e2e', d['e2e']['value'], 'kernel us', d['roofline']['kernel_ms']*1e3, 'frac', d['roofline']['frac'])"
GCA_NO_PDL=1 timeout 900 python bench.py --no-cpu > gpurun_out/bench_nopdl.json 2>> gpurun_out/bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_nopdl.json')); print('NOPDL step us', d['ms_per_step']*1e3, 'value', d['value'], 'e2e', d['e2e']['value'], 'kernel us', d['roofline']['kernel_ms']*1e3, 'frac', d['roofline']['frac'])"
tail -n 5 gpurun_out/bench_n1.err
