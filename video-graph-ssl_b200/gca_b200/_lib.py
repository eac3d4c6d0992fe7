"""ctypes binding of libgca_b200.so (include/gca_b200.h).

There is no CPU fallback anywhere in this package: if the shared library is missing every op raises
`GcaLibraryError` (build it with `python video-graph-ssl_b200/build.py` or `__graft_entry__.build()`),
and if no CUDA device is present the library's compute entry points return GCA_ERR_CUDA.
"""
import ctypes
import os
import threading
from ctypes import c_char_p, c_float, c_int, c_longlong, c_size_t, c_uint, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GCA_B200_LIB") or os.path.join(_HERE, "libgca_b200.so")   # (override: kernel bring-up builds)

GCA_OK = 0
GCA_F32, GCA_BF16 = 0, 1
ALGO = {"auto": 0, "ffma": 1, "tcgen05": 2, "tc32": 3}


GRAPH_THRESHOLD, GRAPH_TOPK, GRAPH_EDGE_DROP, GRAPH_SYMNORM = 1, 2, 4, 8


class GraphOpts(ctypes.Structure):
    """GcaGraphOpts of include/gca_b200.h (default-OFF variants of the graph head)."""
    _fields_ = [("flags", c_uint), ("tau", c_float), ("topk", c_int), ("p_drop", c_float)]


class GcaLibraryError(RuntimeError):
    pass


class GcaError(RuntimeError):
    def __init__(self, fn, code, msg):
        super().__init__("%s failed with code %d: %s" % (fn, code, msg))
        self.code = code


# name -> (restype, argtypes); must list every symbol include/gca_b200.h declares (tests check this)
SIGNATURES = {
    "gca_version": (c_int, []),
    "gca_last_error": (c_char_p, []),
    "gca_sm_count": (c_int, []),
    "gca_launch_count": (c_longlong, []),
    "gca_enqueue_devptr": (c_int, [c_void_p, c_int, c_longlong, c_longlong, c_longlong, c_int, c_void_p, c_int, c_void_p,
                                   c_void_p]),
    "gca_infonce_partials": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_longlong, c_int, c_float, c_int, c_int,
                                     c_void_p, c_size_t, c_void_p]),
    "gca_enqueue": (c_int, [c_void_p, c_int, c_longlong, c_longlong, c_longlong, c_int, c_void_p, c_int, c_longlong,
                            c_void_p]),
    "gca_infonce_workspace_bytes": (c_size_t, [c_int, c_longlong, c_int, c_int, c_int]),
    "gca_infonce_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_longlong, c_int, c_float, c_int,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_size_t, c_void_p]),
    "gca_moco_step": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_longlong, c_int, c_float, c_int,
                              c_void_p, c_int, c_longlong, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_size_t, c_void_p]),
    "gca_infonce_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_longlong, c_int, c_float, c_int,
                                c_void_p, c_float, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gca_infonce_shard_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_longlong, c_int, c_float, c_int,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gca_infonce_shard_combine": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gca_infonce_shard_finish": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float,
                                         c_void_p, c_void_p, c_void_p]),
    "gca_graph_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                              c_float, c_int, c_float, c_uint, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                              c_void_p]),
    "gca_graph_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int,
                              c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int, c_float, c_uint,
                              c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gca_graph_fwd_ex": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                 c_float, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                 c_void_p]),
    "gca_graph_bwd_ex": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int, c_float, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gca_graph_workspace_bytes": (c_size_t, [c_int, c_int]),
    "gca_negcos_fwd_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                   c_void_p]),
    "gca_negcos_workspace_bytes": (c_size_t, [c_int, c_int]),
    "gca_ema_chunk_bytes": (c_size_t, []),
    "gca_ema_update": (c_int, [c_void_p, c_int, c_float, c_float, c_void_p]),
    "gca_keys_exchange_bytes": (c_size_t, [c_int, c_int, c_int]),
    "gca_keys_exchange": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "gca_moco_step_peer": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_longlong, c_int, c_float, c_int,
                                   c_int, c_int, c_void_p, c_void_p, c_int, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_size_t, c_void_p]),
    "gca_shard_peer_bytes": (c_size_t, [c_int, c_int, c_int]),
    "gca_shard_step_peer": (c_int, [c_void_p, c_void_p, c_int, c_int, c_longlong, c_int, c_float, c_int,
                                    c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_size_t, c_void_p]),
    "gca_moco_step_proj": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_longlong, c_int, c_float, c_int,
                                   c_void_p, c_int, c_longlong, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_size_t, c_void_p]),
    "gca_bn1d_fwd": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_float, c_int, c_int, c_void_p, c_void_p,
                             c_void_p, c_void_p, c_void_p, c_void_p]),
    "gca_bn1d_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                             c_void_p, c_void_p, c_void_p, c_void_p]),
    "gca_bank_logits": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_longlong, c_float, c_void_p, c_void_p]),
    "gca_bank_dx_workspace_bytes": (c_size_t, [c_int, c_int]),
    "gca_bank_dx": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_longlong, c_float, c_void_p, c_void_p, c_size_t,
                            c_void_p]),
    "gca_bank_update_workspace_bytes": (c_size_t, [c_int, c_int]),
    "gca_bank_update": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_longlong, c_float, c_float, c_void_p, c_size_t,
                                c_void_p]),
    "gca_workspace_set_done_flag": (c_int, [c_void_p, c_void_p, c_void_p]),
    "gca_plan_begin": (c_int, []),
    "gca_plan_end": (c_int, [c_void_p]),
    "gca_plan_run": (c_int, [c_void_p, c_void_p]),
    "gca_plan_launches": (c_int, [c_void_p]),
    "gca_plan_destroy": (None, [c_void_p]),
    "gca_sim_topk_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "gca_sim_topk": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                             c_size_t, c_void_p]),
}

_lib = None


def load():
    """Load the shared library once; raise loudly when it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GcaLibraryError(
                "%s is missing: build it with `python video-graph-ssl_b200/build.py` "
                "(there is no CPU or PyTorch fallback for the hot path)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error():
    return load().gca_last_error().decode("utf-8", "replace")


def check(fn_name, rc):
    if rc != GCA_OK:
        raise GcaError(fn_name, rc, last_error())


_pending = threading.local()          # device of the first CUDA tensor handed to ptr() since the last call()


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL).  Remembers the tensor's CUDA device for the next `call`."""
    if t is None:
        return None
    if t.is_cuda and getattr(_pending, "dev", None) is None:
        _pending.dev = t.device.index
    return c_void_p(t.data_ptr())


def call(name, *args):
    """Invoke a library entry point and raise on a non-zero code.  The library sizes its launches for the CURRENT device
    (SM count, workspace carve) and launches on it, so the call runs with the device of its tensor arguments current --
    the arguments are built with `ptr()` right before the call, which is where the device is picked up."""
    dev = getattr(_pending, "dev", None)
    _pending.dev = None
    fn = getattr(load(), name)
    if dev is not None:
        import torch
        if torch.cuda.current_device() != dev:
            with torch.cuda.device(dev):
                rc = fn(*args)
            check(name, rc)
            return
    rc = fn(*args)
    check(name, rc)


class LaunchPlan(object):
    """gca_plan_* of include/gca_b200.h: the launches of one step recorded once, re-issued by `run(stream)` with one library
    call (three host-side launch calls for a head step; programmatic dependent launch between them and across steps)."""

    def __init__(self, handle, device):
        self.handle, self.device = handle, device
        lib = load()
        self._run = lib.gca_plan_run
        self.launches = int(lib.gca_plan_launches(handle))

    @classmethod
    def record(cls, issue, device):
        """`issue()` makes the library calls to record (tcgen05-family step entry points only) with `device` current."""
        import torch
        lib = load()
        with torch.cuda.device(device):
            check("gca_plan_begin", lib.gca_plan_begin())
            handle = c_void_p()
            try:
                issue()
            except BaseException:
                lib.gca_plan_end(ctypes.byref(handle))
                if handle.value:
                    lib.gca_plan_destroy(handle)
                raise
            check("gca_plan_end", lib.gca_plan_end(ctypes.byref(handle)))
        return cls(handle, device)

    def run(self, stream):
        """Re-issue the recorded launches on `stream` (raw cudaStream_t as int / c_void_p) of the plan's device."""
        rc = self._run(self.handle, stream)
        if rc != GCA_OK:
            check("gca_plan_run", rc)

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h is not None and _lib is not None:
            try:
                _lib.gca_plan_destroy(h)
            except Exception:
                pass
