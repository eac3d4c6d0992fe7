#!/usr/bin/env python
"""Turns the raw ncu outputs of tools/gpu_r02_evidence.sh (gpurun_out/) into the committed summaries under profiles/:
    r02_launches.txt               per-kernel launch counts / mean durations of the bench command (ncu launch list)
    r02_ncu_full_head_kernels.txt  per-launch metrics of the head kernels from the `ncu --set full` capture"""
import csv, io, json, os, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

def launches():
    rows = [r for r in csv.reader(open(os.path.join(G, "r02_launches.csv"))) if len(r) > 10]
    h = rows[0]
    ki, mi, vi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1; a[1] += float(r[vi].replace(",", ""))
    bench = json.loads(open(os.path.join(G, "r02_bench_n1.json")).read().strip().splitlines()[-1])
    out = ["# r02 (final) -- ncu launch list of `python bench.py --steps 30 --warmup 3 --no-cpu --no-secondary` (tools/gpu_r02_evidence.sh)",
           "# ncu --metrics gpu__time_duration.sum --clock-control none -c 1200; cold-cache, serialised launches: compare SHARES, not absolutes",
           "# kernel | launches | mean ns | total ns"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("%s | %d | %d | %d" % (k[:120], n, t / n, t))
    ours = {k: v for k, v in agg.items() if "gca::" in k}
    per = {k: v[1] / v[0] for k, v in ours.items()}
    tot = sum(per.values())
    out.append("# per step (one launch each of our kernels): " + "; ".join("%s %.1f us (%d%%)" % (k.split("(")[0].replace("void ", "").replace("gca::", ""), v / 1e3, round(100 * v / tot)) for k, v in sorted(per.items(), key=lambda kv: -kv[1])))
    out.append("# bench.py's own figures (the full run of the same call, profiles/r02_bench_n1.json): %.2f us per step back to back (CUDA events), "
               "stream kernel %.2f us per launch in a launch train (%d %% of the step)" % (bench["ms_per_step"] * 1e3, bench["roofline"]["kernel_ms"] * 1e3,
               round(100 * bench["roofline"]["kernel_ms"] / bench["ms_per_step"])))
    out.append("# the at::FillFunctor<unsigned char> launches are bench.py's 256 MiB L2 flush between isolated / e2e iterations (outside the per-step events)")
    open(os.path.join(P, "r02_launches.txt"), "w").write("\n".join(out) + "\n")

def full():
    rep = os.path.join(G, "r02_prof_head.ncu-rep")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h = rows[0]
    want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
            ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
            ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
            ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
            ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots active %"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of max"),
            ("launch__registers_per_thread", "registers/thread"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
            ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__shared_mem_per_block_dynamic", "dynamic smem/block")]
    ki = h.index("Kernel Name")
    units = rows[1]
    out = ["# r02 (final) -- `ncu --set full --clock-control none --import-source on -k regex:infonce_tcx|infonce_finalize|infonce_prep -s 60 -c 6`",
           "# on `python bench.py --steps 30 --warmup 3 --no-cpu --no-secondary` (tools/gpu_r02_evidence.sh; the same command exited 0 without ncu first).",
           "# Per-launch values; ncu replays each launch with flushed caches, so durations are cold and serialised (compare with profiles/r02_launches.txt).", ""]
    for r in rows[2:]:
        out.append(r[ki].split("(")[0].replace("void ", "").replace("gca::", ""))
        for m, label in want:
            if m in h:
                i = h.index(m)
                out.append("    %-28s %s %s" % (label, r[i], units[i]))
        out.append("")
    open(os.path.join(P, "r02_ncu_full_head_kernels.txt"), "w").write("\n".join(out))
    # dram traffic of the dominant kernel for bench.py's roofline.traffic
    for r in rows[2:]:
        if "infonce_tcx" in r[ki]:
            rd, wr = float(r[h.index("dram__bytes_read.sum")].replace(",", "")), float(r[h.index("dram__bytes_write.sum")].replace(",", ""))
            print("tcx dram read %s %s write %s" % (rd, units[h.index("dram__bytes_read.sum")], wr))

launches()
full()
