cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 120 python tools/tc_bringup.py 256 65536 2>&1 | tail -n 3
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print('FUSED step us', d['ms_per_step']*1e3, 'value', d['value'], 'e2e', d['e2e']['value'], 'kernel us', d['roofline']['kernel_ms']*1e3, 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'])"
GCA_NO_FUSE=1 timeout 300 python bench.py --no-cpu > gpurun_out/bench_nofuse.json 2>> gpurun_out/bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_nofuse.json')); print('2-KERNEL step us', d['ms_per_step']*1e3, 'value', d['value'], 'e2e', d['e2e']['value'])"
tail -n 5 gpurun_out/bench_n1.err
