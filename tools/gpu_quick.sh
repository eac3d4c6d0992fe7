#!/bin/bash
# where do finalize's partial reads come from?  one-pass ncu (no cache flush, no replay): DRAM bytes of the step's kernels
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
timeout 200 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
  -k regex:"infonce_tcx|infonce_finalize|infonce_prep" -s 90 -c 9 --csv --log-file gpurun_out/r02_step_dram_nocachectl.csv \
  python bench.py --steps 40 --warmup 3 --no-cpu --no-secondary > gpurun_out/ncu_nocc.log 2>&1
echo "exit $?"
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r02_step_dram_nocachectl.csv")) if len(r) > 10]
h = rows[0]; ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
d = {}
for r in rows[1:]:
    d.setdefault((int(r[ii]), r[ki].split("(")[0][-40:]), {})[r[mi]] = r[vi]
for k, v in sorted(d.items()):
    print(k, v)
PY
