#!/bin/bash
# stream-kernel iteration: parity subset, kernel/step timing for every lib variant given, timeline of the default build
timeout 600 python -m pytest tests -m gpu -q -x -k "not multi_gpu" > gpurun_out/r2_pytest_k.log 2>&1; tail -4 gpurun_out/r2_pytest_k.log
timeout 200 python tools/tcx_time.py 2>&1 | tail -2
for v in "$@"; do
  GCA_B200_LIB=$PWD/video-graph-ssl_b200/build/variants/lib_$v.so timeout 200 python tools/tcx_time.py 2>&1 | tail -2
done
timeout 120 python tools/tc_timeline.py > gpurun_out/r2_tl_k.txt 2>&1; grep -v "^   mma i=3: wait" gpurun_out/r2_tl_k.txt | tail -24
