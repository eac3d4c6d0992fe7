"""Momentum encoder update as one kernel launch (gca_ema_update).

Mirror of `Trainer._momentum_update(model, model_ema, m)` (tools/train_video_contrast_dis.py:176-180):
    for p1, p2 in zip(model.parameters(), model_ema.parameters()):  p2.data.mul_(m).add_(p1.detach().data, alpha=1 - m)
"""
import ctypes
import weakref

import numpy as np
import torch

from . import _lib
from ._lib import ptr

CHUNK = 16384          # elements per descriptor: 64 KB of fp32, one CTA's worth of work


def _dense(t):
    """Every element of the storage span is used exactly once (any permutation of a contiguous layout)."""
    expect = 1
    for size, stride in sorted(((sz, st) for sz, st in zip(t.shape, t.stride()) if sz > 1), key=lambda x: x[1]):
        if stride != expect:
            return False
        expect *= size
    return True


class MomentumUpdater(object):
    """Built once for a (model, model_ema) pair; `step(m)` updates every parameter of model_ema in a single launch.
    Parameters must be fp32, dense with the same layout on both sides, on one CUDA device, and keep their storage (rebuild after re-allocation)."""

    def __init__(self, model, model_ema):
        pairs = list(zip(model.parameters(), model_ema.parameters()))
        if not pairs:
            raise ValueError("no parameters")
        rec = []
        dev = pairs[0][0].device
        for p, e in pairs:
            if not (p.is_cuda and e.is_cuda):
                raise RuntimeError("MomentumUpdater needs CUDA parameters; there is no CPU path")
            if p.dtype != torch.float32 or e.dtype != torch.float32 or p.shape != e.shape:
                raise TypeError("MomentumUpdater handles matching fp32 parameters")
            if p.stride() != e.stride() or not _dense(p):                 # channels_last(_3d) weights are fine
                raise ValueError("parameter pairs must be dense and share one memory layout")
            n = p.numel()
            for off in range(0, n, CHUNK):
                rec.append((e.data_ptr() + 4 * off, p.data_ptr() + 4 * off, min(CHUNK, n - off)))
        assert int(_lib.load().gca_ema_chunk_bytes()) == 24
        table = np.array(rec, dtype=np.int64).reshape(-1, 3)
        self.table = torch.from_numpy(table).to(dev)                  # [nchunks, 3] int64 == {float*, const float*, long long}
        self.nchunks = table.shape[0]
        self.device = dev
        self.numel = sum(p.numel() for p, _ in pairs)
        self._keep = pairs                                            # keeps the storages (and their addresses) alive

    def step(self, m):
        # alpha = 1 - m is formed in double and rounded once, like the reference's `add_(p1, alpha=1 - m)` (train...:179)
        _lib.call("gca_ema_update", ptr(self.table), self.nchunks, float(m), float(1.0 - float(m)),
                  ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))


_UPDATERS = weakref.WeakKeyDictionary()          # model_ema -> (weakref to model, MomentumUpdater); dies with model_ema


def momentum_update(model, model_ema, m):
    """Drop-in for `Trainer._momentum_update(model, model_ema, m)`.  The descriptor table is built on first use and kept
    per `model_ema` (rebuilt if it is later paired with another model or a parameter was re-allocated)."""
    slot = _UPDATERS.get(model_ema)
    up = None
    if slot is not None:
        ref, cand = slot
        if ref() is model and all(p.data_ptr() == q.data_ptr() and e.data_ptr() == f.data_ptr()
                                  for (p, e), (q, f) in zip(zip(model.parameters(), model_ema.parameters()), cand._keep)):
            up = cand
    if up is None:
        up = MomentumUpdater(model, model_ema)
        _UPDATERS[model_ema] = (weakref.ref(model), up)
    up.step(m)
