"""CPU: the C-ABI library builds, loads and exports every symbol include/gca_b200.h declares; argument validation
returns error codes (never aborts); no compute call is made without a GPU."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "gca_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gca_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound(lib):
    from gca_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 17
    for s in syms:
        assert hasattr(lib, s), "libgca_b200.so does not export %s" % s
        assert s in _lib.SIGNATURES, "ctypes binding misses %s" % s
    assert sorted(_lib.SIGNATURES) == syms, "binding lists symbols the header does not declare"
    assert lib.gca_version() == 1


def test_library_is_sm100a_with_tensor_core_path():
    so = os.path.join(ROOT, "video-graph-ssl_b200", "gca_b200", "libgca_b200.so")
    if not os.path.exists(so):
        pytest.skip("library not built")
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    elf = subprocess.run([cuobjdump, "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in elf
    sass = subprocess.run([cuobjdump, "-sass", so], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "STTM"):     # tcgen05.mma, TMA, tcgen05.ld / st
        assert mnemonic in sass, mnemonic


def test_argument_validation_without_gpu(lib):
    from gca_b200 import _lib
    null = ctypes.c_void_p(0)
    one = ctypes.c_void_p(16)
    # enqueue: bad dtype / shard range / too many rows / pointer outside ring
    assert lib.gca_enqueue(one, 7, 16, 0, 16, 8, one, 1, 0, null) == -1
    assert b"dtype" in lib.gca_last_error()
    assert lib.gca_enqueue(one, 0, 16, 8, 4, 8, one, 1, 0, null) == -1
    assert lib.gca_enqueue(one, 0, 16, 0, 16, 8, one, 17, 0, null) == -1            # N > K: undefined upstream
    assert lib.gca_enqueue(one, 0, 16, 0, 16, 8, one, 1, 16, null) == -1
    assert lib.gca_enqueue(one, 0, 16, 0, 16, 6, one, 1, 0, null) == -1             # d % 4
    assert lib.gca_enqueue(null, 0, 16, 0, 16, 8, one, 1, 0, null) == -1
    assert lib.gca_enqueue(one, 0, 16, 0, 16, 8, one, 0, 0, null) == 0              # empty enqueue is a no-op
    # infonce: unsupported shapes are reported, not silently rerouted
    args = [one, one, one, 1, 4, 64, 100, 1.0, 1] + [one] * 8 + [one, 1 << 20, null]
    assert lib.gca_infonce_fwd(*args) == -2                                         # ffma: d % 32
    args[6] = 64
    args[8] = 2
    assert lib.gca_infonce_fwd(*args) == -2                                         # tcgen05: d != 128
    args[7] = 0.0
    assert lib.gca_infonce_fwd(*args) == -1                                         # inv_T
    # graph: T range and flags
    g = [one, one, 4, 1, one, 8, 1, 40, 2, one, 0.5, 3, 1.0, 0, one, one, one, one, one, 1 << 20, null]
    assert lib.gca_graph_fwd(*g) == -1
    g[7] = 4
    g[13] = 5
    assert lib.gca_graph_fwd(*g) == -2
    # top-k: k range
    assert lib.gca_sim_topk(one, one, 4, 8, 16, 65, 1, one, null, one, 1 << 30, null) == -1
    assert lib.gca_sim_topk(one, one, 4, 8, 16, 9, 1, one, null, one, 1 << 30, null) == -1
    assert lib.gca_sim_topk(one, one, 4, 8, 16, 4, 1, one, null, one, 8, null) == -4  # workspace too small


def test_launch_plan_entry_points_without_gpu(lib):
    """gca_plan_*: misuse returns error codes; without a CUDA device recording cannot start (no CPU path)."""
    h = ctypes.c_void_p()
    assert lib.gca_plan_end(ctypes.byref(h)) == -1                                  # not recording
    assert lib.gca_plan_run(None, None) == -1
    assert lib.gca_plan_launches(None) == 0
    lib.gca_plan_destroy(None)
    if not torch.cuda.is_available():
        assert lib.gca_plan_begin() == -3                                           # GCA_ERR_CUDA
        assert b"no CUDA device" in lib.gca_last_error()


def test_workspace_sizes(lib):
    assert lib.gca_infonce_workspace_bytes(256, 65536, 128, 1, 0) >= 74 * 256 * 128 * 4
    assert lib.gca_infonce_workspace_bytes(0, 65536, 128, 1, 0) == 0
    assert lib.gca_graph_workspace_bytes(128, 8) >= 128 * 64 * 4
    assert lib.gca_negcos_workspace_bytes(128, 1024) >= 128 * 4
    assert lib.gca_sim_topk_workspace_bytes(10, 20, 8, 5) >= 10 * 20 * 4


def test_no_cpu_fallback():
    """CPU tensors are rejected loudly; nothing in the product imports the oracle."""
    import gca_b200
    moco = gca_b200.RGBMoCo(32, K=64)
    with pytest.raises(RuntimeError, match="no CPU path"):
        moco(torch.randn(4, 32), torch.randn(4, 32))
    with pytest.raises(RuntimeError, match="no CPU path"):
        gca_b200.D()(torch.randn(4, 8), torch.randn(4, 8))
    with pytest.raises(RuntimeError, match="no CPU path"):
        gca_b200.TemporalGraphAug(8, sub_sample=False)(torch.randn(2, 8, 4, 1, 1))
    pkg = os.path.join(ROOT, "video-graph-ssl_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f


def test_missing_library_fails_loudly(monkeypatch):
    from gca_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libgca_b200.so")
    with pytest.raises(_lib.GcaLibraryError, match="no CPU or PyTorch fallback"):
        _lib.load()
