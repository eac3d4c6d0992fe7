#!/usr/bin/env python
"""Build libgca_b200.so (sm_100a only) in-tree with nvcc: one object per .cu in parallel, then one link.

    python video-graph-ssl_b200/build.py [--force] [--verbose]

The .so lands in video-graph-ssl_b200/gca_b200/ (git-ignored, shipped to the GPU box with the snapshot).
"""
import argparse
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "gca_b200", "libgca_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "--use_fast_math", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
         "--expt-relaxed-constexpr"]
# accuracy-sensitive translation units keep IEEE expf/logf/div (the graph head must match torch to ~1e-6)
PRECISE = {"graph.cu", "negcos.cu", "finalize.cu", "sim_tc.cu", "bank.cu"}


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def newest_input():
    t = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            t = max(t, os.path.getmtime(os.path.join(root, f)))
    return max(t, os.path.getmtime(os.path.abspath(__file__)))


def compile_one(src, verbose):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    flags = [f for f in FLAGS if not (f == "--use_fast_math" and src in PRECISE)]
    cmd = [nvcc()] + ARCH + flags + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(os.path.join(OBJ, src[:-3] + ".ptxas.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (src, log))
    if verbose:
        print(log)
    return obj


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= newest_input():
        return LIB
    srcs = sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: compile_one(s, verbose), srcs))
    cmd = [nvcc()] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
