set -x
nvidia-smi -L
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 120 python tools/tc_bringup.py 128 128 > gpurun_out/tc1.log 2>&1; echo "tc1 exit $?" >> gpurun_out/tc1.log
timeout 120 python tools/tc_bringup.py 128 1024 > gpurun_out/tc2.log 2>&1; echo "tc2 exit $?" >> gpurun_out/tc2.log
timeout 120 python tools/tc_bringup.py 256 65536 > gpurun_out/tc3.log 2>&1; echo "tc3 exit $?" >> gpurun_out/tc3.log
if ! grep -q "err|=[0-9.]*e-0[4-9]" gpurun_out/tc1.log; then timeout 600 python tools/tc_bringup.py --sweep > gpurun_out/tc_sweep.log 2>&1; fi
timeout 1500 python -m pytest tests -m gpu -q -k "not tcgen05 and not bf16 and not headline and not shard_abi" > gpurun_out/pytest_a.log 2>&1; echo "exit $?" >> gpurun_out/pytest_a.log
timeout 1500 python -m pytest tests -m gpu -q -k "tcgen05 or bf16 or headline or shard_abi" > gpurun_out/pytest_b.log 2>&1; echo "exit $?" >> gpurun_out/pytest_b.log
tail -5 gpurun_out/tc1.log gpurun_out/tc2.log gpurun_out/tc3.log
tail -30 gpurun_out/pytest_a.log
tail -30 gpurun_out/pytest_b.log
