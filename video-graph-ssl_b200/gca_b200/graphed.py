"""One trainer-style head step captured as a CUDA graph.

At the headline shape the whole head (logits + loss + gradient + top-k + enqueue) is ~10 us of GPU work, less than
the cost of launching its three kernels from Python one by one (SURVEY.md section 7.2 "microsecond budgets").
`GraphedMoCoStep` captures

    gca_moco_step = queue-streaming kernel + fixed-order finalize (loss, lse, rank, top-1/top-5 hits, d loss/d q) with
                    the in-place ring enqueue riding in the finalize launch; the ring pointer lives in device memory

once, over static input/output buffers, and replays it per step -- the same arithmetic as
`RGBMoCo.forward` -> `NCESoftmaxLoss` -> `loss.backward()` -> `accuracy` (train_video_contrast_dis.py:411-428)
with grad_output = 1.  The python-side `moco.index` is advanced in lock-step so the module stays consistent.
"""
import ctypes

import torch

from . import _lib
from . import functional as GF
from ._lib import ptr


class GraphedMoCoStep(object):
    def __init__(self, moco, batch, n_enqueue=None, algo=None, state=None):
        mem = moco.memory
        if not mem.is_cuda:
            raise RuntimeError("GraphedMoCoStep needs the queue on a CUDA device; there is no CPU path")
        self.moco = moco
        self.B, self.N = int(batch), int(batch if n_enqueue is None else n_enqueue)
        self.K, self.d = mem.shape
        self.algo = moco.algo if algo is None else algo
        dev = mem.device
        # one packed input buffer so an end-to-end caller needs a single host->device copy per step
        self.inputs = torch.zeros(2 * self.B + self.N, self.d, dtype=torch.float32, device=dev)
        self.q, self.k = self.inputs[:self.B], self.inputs[self.B:2 * self.B]
        self.all_k = self.inputs[2 * self.B:]
        # one packed output buffer: [loss, top1_hits, top5_hits, pad] + dq[B, d] -> a single device->host copy
        self.outputs = torch.zeros(4 + self.B * self.d, dtype=torch.float32, device=dev)
        self.loss = self.outputs[0:1]
        self.hits = self.outputs[1:3].view(torch.int32)
        self.dq = self.outputs[4:].view(self.B, self.d)
        self.loss_rows = torch.empty(self.B, dtype=torch.float32, device=dev)
        self.lse = torch.empty_like(self.loss_rows)
        self.pos = torch.empty_like(self.loss_rows)
        self.rank = torch.empty(self.B, dtype=torch.int32, device=dev)
        # ring pointer + ticket in device memory; several captured steps over the same queue share one state tensor
        self.state = state if state is not None else torch.tensor([moco.index, 0], dtype=torch.int64, device=dev)
        self.qd = GF.queue_dtype_code(mem)
        nbytes = GF.infonce_workspace_bytes(self.B, self.K, self.d, self.qd, self.algo)
        self.ws = torch.zeros(nbytes, dtype=torch.uint8, device=dev)        # private: graphs must not share scratch
        self.graph = None
        self.launches_per_step = 0

    def _enqueue_work(self, stream):
        m = self.moco
        _lib.call("gca_moco_step", ptr(self.q), ptr(self.k), ptr(m.memory), self.qd, self.B, self.K, self.d, 1.0 / m.T,
                  _lib.ALGO[self.algo], ptr(self.all_k), self.N, 0, ptr(self.state), None,
                  ptr(self.loss), ptr(self.loss_rows), ptr(self.lse), ptr(self.pos), ptr(self.rank), ptr(self.hits),
                  ptr(self.dq), ptr(self.ws), self.ws.numel(), stream)

    def capture(self):
        lib = _lib.load()
        dev = self.moco.memory.device
        # un-captured warm-up (loads the kernels, sets their attributes); the rows it overwrites are restored
        slots = (torch.arange(self.N, device=dev) + self.moco.index) % self.K
        saved = self.moco.memory[slots].clone()
        self.state.copy_(torch.tensor([self.moco.index, 0], dtype=torch.int64))
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            self._enqueue_work(ctypes.c_void_p(side.cuda_stream))
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.moco.memory[slots] = saved
        self.state.copy_(torch.tensor([self.moco.index, 0], dtype=torch.int64))
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        n0 = lib.gca_launch_count()
        with torch.cuda.graph(g):
            self._enqueue_work(ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        self.launches_per_step = int(lib.gca_launch_count() - n0)
        self.graph = g
        return self

    def step(self, q=None, k=None, all_k=None):
        """Replay one step.  Tensors given here are copied into the static buffers first (device-side copies);
        results are in .loss, .hits (int32 [2]), .dq, .rank, .lse until the next replay."""
        if self.graph is None:
            self.capture()
        if q is not None:
            self.q.copy_(q, non_blocking=True)
        if k is not None:
            self.k.copy_(k, non_blocking=True)
            if all_k is None and self.N == self.B:
                self.all_k.copy_(k, non_blocking=True)
        if all_k is not None:
            self.all_k.copy_(all_k, non_blocking=True)
        self.graph.replay()
        self.moco.index = (self.moco.index + self.N) % self.K
        return self.loss


class GraphedReplicaStep(GraphedMoCoStep):
    """The same captured step for data-parallel replicas (the reference's own parallelisation: a replicated queue that
    every rank updates with the keys of ALL ranks, train_video_contrast_dis.py:222 + mem_moco.py:81-83).

    The key all-gather (NCCL over NVLink) is captured on a side stream and overlaps the queue-streaming kernel, which
    only needs the LOCAL keys for its positives; the gathered keys are first needed by the enqueue at the end of the step.
    Every rank must hold an identical queue and ring pointer (same seed or a broadcast, as upstream)."""

    def __init__(self, moco, batch, group=None, algo=None, state=None):
        import torch.distributed as dist
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        super(GraphedReplicaStep, self).__init__(moco, batch, n_enqueue=batch * self.world, algo=algo, state=state)
        self.side = torch.cuda.Stream(moco.memory.device)
        self.keys_ready = torch.cuda.Event()

    def _enqueue_work(self, stream):
        import torch.distributed as dist
        m = self.moco
        dev = m.memory.device
        main = torch.cuda.current_stream(dev)
        self.side.wait_stream(main)
        with torch.cuda.stream(self.side):
            dist.all_gather_into_tensor(self.all_k, self.k, group=self.group)
            self.keys_ready.record(self.side)
        # one C call: prep + queue sweep, then (after the event) finalize with the enqueue of the gathered keys riding in it
        _lib.call("gca_moco_step", ptr(self.q), ptr(self.k), ptr(m.memory), self.qd, self.B, self.K, self.d, 1.0 / m.T,
                  _lib.ALGO[self.algo], ptr(self.all_k), self.N, 0, ptr(self.state), ctypes.c_void_p(self.keys_ready.cuda_event),
                  ptr(self.loss), ptr(self.loss_rows), ptr(self.lse), ptr(self.pos), ptr(self.rank), ptr(self.hits),
                  ptr(self.dq), ptr(self.ws), self.ws.numel(), stream)
        main.wait_stream(self.side)
