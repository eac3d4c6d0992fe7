"""One trainer-style head step captured once (launch plan / CUDA graph) and re-issued per step.

At the headline shape the whole head (logits + loss + gradient + top-k + enqueue) is ~10 us of GPU work, less than
the cost of launching its three kernels from Python one by one (SURVEY.md section 7.2 "microsecond budgets").
`GraphedMoCoStep` captures

    gca_moco_step = queue-streaming kernel + fixed-order finalize (loss, lse, rank, top-1/top-5 hits, d loss/d q) with
                    the in-place ring enqueue riding in the finalize launch; the ring pointer lives in device memory

once, over static input/output buffers, and replays it per step -- the same arithmetic as
`RGBMoCo.forward` -> `NCESoftmaxLoss` -> `loss.backward()` -> `accuracy` (train_video_contrast_dis.py:411-428)
with grad_output = 1.  The python-side `moco.index` is advanced in lock-step so the module stays consistent.

`step()` re-issues the captured work from a launch plan when the step is recordable (tcgen05 family: bf16 queue, d == 128;
include/gca_b200.h, gca_plan_*): the library's own three launches per step, with programmatic dependent launch between
them and across consecutive steps -- 21.3 us per step against 24.2 us for one CUDA-graph launch per step at the headline
shape, and less host time per step.  The CUDA graph of the same work stays available (`.graph`, `prefer_graph=True`).
"""
import ctypes

import torch

from . import _lib
from . import functional as GF
from ._lib import ptr


class GraphedMoCoStep(object):
    def __init__(self, moco, batch, n_enqueue=None, algo=None, state=None, want_rank=True):
        """want_rank=False: only the top-1 / top-5 hit counts are produced (what `accuracy(output, labels, topk=(1, 5))`
        reports, train_video_contrast_dis.py:428), not the per-row rank of the positive: the kernel may then stop counting
        once every row of a warp is past rank 8 (include/gca_b200.h, GCA_TOPK_RANK_CAP); .rank is None."""
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
        self.rank = torch.empty(self.B, dtype=torch.int32, device=dev) if want_rank else None
        # ring pointer + ticket in device memory; several captured steps over the same queue share one state tensor
        self.state = state if state is not None else torch.tensor([moco.index, 0], dtype=torch.int64, device=dev)
        self.qd = GF.queue_dtype_code(mem)
        nbytes = GF.infonce_workspace_bytes(self.B, self.K, self.d, self.qd, self.algo)
        self.ws = torch.zeros(nbytes, dtype=torch.uint8, device=dev)        # private: graphs must not share scratch
        self.graph = None
        self.plan = None                      # _lib.LaunchPlan of the same work (preferred by step() unless prefer_graph)
        self._io_done = None                  # numpy view of the host completion word (capture_host_io with zero-copy outputs)
        self._io_seq = 0
        self.prefer_graph = False
        self.launches_per_step = 0

    def _plannable(self):
        """The step's launches can be recorded into a launch plan: every one of them is this library's own (tcgen05 family)."""
        return self.qd == _lib.GCA_BF16 and self.d == 128 and self.algo in ("auto", "tcgen05")

    def _raw_stream(self):
        dev = self.moco.memory.device
        return torch._C._cuda_getCurrentRawStream(dev.index if dev.index is not None else torch.cuda.current_device())

    def _run_plan(self, plan):
        dev = self.moco.memory.device.index
        if dev is not None and torch.cuda.current_device() != dev:
            with torch.cuda.device(dev):
                plan.run(self._raw_stream())
        else:
            plan.run(self._raw_stream())

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
        if self._plannable():
            self.plan = _lib.LaunchPlan.record(lambda: self._enqueue_work(None), dev)
        return self

    def capture_host_io(self, host_in, host_out, zero_copy_out=True, zero_copy_in=False):
        """A second graph for callers whose step inputs live in (pinned) host memory: H2D copy of `host_in` into the packed
        input buffer -> the step -> the packed outputs (loss, top-1/top-5 hits, dq) in `host_out`, all in ONE graph launch.
        With zero_copy_out the finalize kernel stores loss | hits | dq straight into the pinned host buffer over PCIe (posted
        writes, complete when the stream is synchronised) instead of a D2H copy node behind it; .loss/.hits/.dq then hold
        the results of step() replays only.  With zero_copy_in there is no H2D copy node either: the first kernel reads q and
        k from the pinned host buffer over PCIe (each exactly once; k is staged on the device for the last kernel) and the
        enqueue CTAs read the new keys from it.  `host_in` / `host_out` are fixed pinned tensors shaped like .inputs / .outputs;
        `step_host_io()` replays the graph (the caller synchronises the stream before reading host_out)."""
        if self.graph is None:
            self.capture()
        if not (host_in.is_pinned() and host_out.is_pinned()):
            raise ValueError("host buffers must be pinned")
        # host_in holds q | k | all_k like .inputs, or only q | k when the enqueue keys come from elsewhere (replica steps)
        full = host_in.shape == self.inputs.shape
        if not (full or host_in.shape == (2 * self.B, self.d)) or host_out.shape != self.outputs.shape:
            raise ValueError("host buffers must match .inputs %r (or its q|k part) and .outputs %r"
                             % (tuple(self.inputs.shape), tuple(self.outputs.shape)))
        dev = self.moco.memory.device
        self._host_io = (host_in, host_out)
        saved = (self.loss, self.hits, self.dq)
        saved_in = (self.q, self.k, self.all_k)
        if zero_copy_in:
            if self.qd != _lib.GCA_BF16 or self.d != 128:
                raise ValueError("zero_copy_in needs the tcgen05 family (bf16 queue, d == 128): its first kernel stages k")
            self.q, self.k = host_in[:self.B], host_in[self.B:2 * self.B]
            if full:
                self.all_k = host_in[2 * self.B:]
        if zero_copy_out:                                            # pinned memory is device-addressable (unified addressing)
            self.loss = host_out[0:1]
            self.hits = host_out[1:3].view(torch.int32)
            self.dq = host_out[4:].view(self.B, self.d)
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                if not zero_copy_in:
                    self.inputs[:host_in.shape[0]].copy_(host_in, non_blocking=True)
                self._enqueue_work(ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
                if not zero_copy_out:
                    host_out.copy_(self.outputs, non_blocking=True)
        finally:
            self.loss, self.hits, self.dq = saved
            self.q, self.k, self.all_k = saved_in
        self.graph_io = g
        # completion word (gca_workspace_set_done_flag): the pad word of the packed output block in host memory receives the
        # number of steps completed on this workspace; step_host_io(wait=True) polls it instead of synchronising the stream
        self._io_done = None
        if zero_copy_out:
            torch.cuda.synchronize(dev)
            word = host_out[3:4].view(torch.int32)
            word.zero_()
            _lib.call("gca_workspace_set_done_flag", ptr(self.ws), ctypes.c_void_p(word.data_ptr()),
                      ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
            self._io_seq = int(self.ws.view(torch.int32)[15]) & 0xffffffff
            self._io_done = host_out.numpy().view("uint32")
        return self

    def step_host_io(self, wait=False):
        """One step on the host buffers given to capture_host_io().  wait=True returns when the results are in `host_out` and
        `host_in` has been consumed: with zero-copy outputs by polling the step's completion word in host memory (about a
        microsecond after the last store), otherwise by synchronising the stream."""
        # a caller that synchronises after every step wants ONE submission: a lone step issued as three launches leaves the
        # GPU waiting for the host between them (measured 49.7 us against 44.2 us per synchronous step)
        self.graph_io.replay()
        if self._io_done is not None:
            self._io_seq = (self._io_seq + 1) & 0xffffffff
        if wait:
            if self._io_done is None:
                torch.cuda.current_stream(self.moco.memory.device).synchronize()
            else:
                w, seq, spins = self._io_done, self._io_seq, 0
                while int(w[3]) != seq:
                    spins += 1
                    if spins > 2000000:                              # ~0.2 s: something is wrong -- fall back to the driver
                        torch.cuda.synchronize(self.moco.memory.device)
                        if int(w[3]) != seq:
                            raise RuntimeError("step_host_io: completion word %d, expected %d" % (int(w[3]), seq))
        self.moco.index = (self.moco.index + self.N) % self.K

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
        if self.plan is not None and not self.prefer_graph:
            self._run_plan(self.plan)
        else:
            self.graph.replay()
        if self._io_done is not None:                                # (same workspace: these steps count as well)
            self._io_seq = (self._io_seq + 1) & 0xffffffff
        self.moco.index = (self.moco.index + self.N) % self.K
        return self.loss


class GraphedReplicaStep(GraphedMoCoStep):
    """The same captured step for data-parallel replicas (the reference's own parallelisation: a replicated queue that
    every rank updates with the keys of ALL ranks, train_video_contrast_dis.py:222 + mem_moco.py:81-83).

    The key all-gather (NCCL over NVLink) is captured on a side stream and overlaps the queue-streaming kernel, which
    only needs the LOCAL keys for its positives; the gathered keys are first needed by the enqueue at the end of the step.
    Every rank must hold an identical queue and ring pointer (same seed or a broadcast, as upstream)."""

    def __init__(self, moco, batch, group=None, algo=None, state=None, exchange=None, fuse_exchange=True, want_rank=True):
        """`exchange`: None -> NCCL all-gather on a side stream.  A gca_b200.peer.PeerKeyExchange (shared by every
        captured step of this process; the steps then must replay in the same order on all ranks) -> the keys travel
        through peer memory: fused into the step's own launches (gca_moco_step_peer: pushed by the first launch,
        enqueued straight from the mailbox by the last; `all_k` is not materialised) or, with fuse_exchange=False, as
        one stand-alone exchange kernel on the side stream."""
        import torch.distributed as dist
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        super(GraphedReplicaStep, self).__init__(moco, batch, n_enqueue=batch * self.world, algo=algo, state=state,
                                                 want_rank=want_rank)
        self.side = torch.cuda.Stream(moco.memory.device)
        self.keys_ready = torch.cuda.Event()
        self.exchange = exchange
        self.fuse_exchange = bool(fuse_exchange) and exchange is not None
        # a peer that misses the exchange timeout leaves this rank's queue un-updated (the kernel skips the enqueue and
        # raises a sticky device flag): poll it every `check_every` replays (one small device->host read) and in check()
        self.check_every = 256
        self._replays = 0

    def _plannable(self):
        return self.fuse_exchange and super(GraphedReplicaStep, self)._plannable()

    def check(self):
        """Raise if a peer ever missed the key-exchange timeout (the replicas' queues have diverged since)."""
        if self.exchange is not None:
            self.exchange.check()

    def step(self, q=None, k=None, all_k=None):
        loss = super(GraphedReplicaStep, self).step(q, k, all_k)
        self._replays += 1
        if self.exchange is not None and self.exchange.timeout_ms > 0 and self._replays % self.check_every == 0:
            self.exchange.check()
        return loss

    def _enqueue_work(self, stream):
        import torch.distributed as dist
        m = self.moco
        dev = m.memory.device
        if self.fuse_exchange:
            x = self.exchange
            _lib.call("gca_moco_step_peer", ptr(self.q), ptr(self.k), ptr(m.memory), self.qd, self.B, self.K, self.d,
                      1.0 / m.T, _lib.ALGO[self.algo], x.W, x.r, ptr(x.table), ptr(x.xstate), x.timeout_ms, ptr(self.state),
                      ptr(self.loss), ptr(self.loss_rows), ptr(self.lse), ptr(self.pos), ptr(self.rank), ptr(self.hits),
                      ptr(self.dq), ptr(self.ws), self.ws.numel(), stream)
            return
        main = torch.cuda.current_stream(dev)
        self.side.wait_stream(main)
        with torch.cuda.stream(self.side):
            if self.exchange is not None:
                self.exchange(self.k, self.all_k, ctypes.c_void_p(self.side.cuda_stream))
            else:
                dist.all_gather_into_tensor(self.all_k, self.k, group=self.group)
            self.keys_ready.record(self.side)
        # one C call: prep + queue sweep, then (after the event) finalize with the enqueue of the gathered keys riding in it
        _lib.call("gca_moco_step", ptr(self.q), ptr(self.k), ptr(m.memory), self.qd, self.B, self.K, self.d, 1.0 / m.T,
                  _lib.ALGO[self.algo], ptr(self.all_k), self.N, 0, ptr(self.state), ctypes.c_void_p(self.keys_ready.cuda_event),
                  ptr(self.loss), ptr(self.loss_rows), ptr(self.lse), ptr(self.pos), ptr(self.rank), ptr(self.hits),
                  ptr(self.dq), ptr(self.ws), self.ws.numel(), stream)
        main.wait_stream(self.side)


class GraphedShardedStep(object):
    """One K-sharded head step (gca_b200.dist.ShardedRGBMoCo, BASELINE config 4) as a single CUDA graph per rank:

        all-gather [q; k]  ->  shard sweep (prep + queue-streaming kernel + partial merge)  ->  all-gather of the
        (max, sum, count) partials  ->  combine  ->  reduce-scatter of the gradient accumulator  ->  finish
        ->  sharded enqueue of the gathered keys (device-resident ring pointer)

    The three NCCL collectives are captured with the kernels, so a step costs one graph launch instead of ~25 python-
    issued operations.  Same arithmetic as ShardedRGBMoCo.forward + NCESoftmaxLoss + backward with grad_output = 1;
    results for the LOCAL rows: .loss (mean over local rows), .dq, .rank, .lse, .loss_rows."""

    def __init__(self, moco, batch_local, algo=None, link=None):
        """`link`: a gca_b200.peer.PeerShardLink (shared by every captured step of this process) -> the q|k gather and the
        cross-rank merge travel through NVLink peer memory inside the step's own launches (gca_shard_step_peer: 4 launches,
        no collective call) instead of three captured NCCL collectives."""
        import torch.distributed as dist
        mem = moco.memory
        if not mem.is_cuda:
            raise RuntimeError("GraphedShardedStep needs the shard on a CUDA device; there is no CPU path")
        self.moco, self.group = moco, moco.group
        self.link = link
        self.W, self.r = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.Bl, self.Bg = int(batch_local), int(batch_local) * self.W
        self.Ks, self.d = mem.shape
        self.algo = moco.algo if algo is None else algo
        dev = mem.device
        f32 = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        self.inputs = f32(2, self.Bl, self.d)                       # [q_loc; k_loc], one H2D copy
        self.q, self.k = self.inputs[0], self.inputs[1]
        self.gathered = f32(self.W, 2, self.Bl, self.d)
        self.qk_all = f32(2, self.Bg, self.d)                       # q_all, k_all (rank-major rows)
        self.pos = f32(self.Bg)
        self.stats = f32(3, self.Bg)
        self.all_stats = f32(self.W, 3, self.Bg)
        self.st = f32(3, self.W, self.Bg)
        self.acc = f32(self.Bg, self.d)
        self.acc_loc = f32(self.Bl, self.d)
        self.lse_all, self.loss_rows_all = f32(self.Bg), f32(self.Bg)
        self.rank_all = torch.zeros(self.Bg, dtype=torch.int32, device=dev)
        self.outputs = f32(4 + self.Bl * self.d)                    # [loss, pad x3] + dq -> one D2H copy
        self.loss = self.outputs[0:1]
        self.dq = self.outputs[4:].view(self.Bl, self.d)
        sl = slice(self.r * self.Bl, (self.r + 1) * self.Bl)
        self.lse, self.loss_rows, self.rank = self.lse_all[sl], self.loss_rows_all[sl], self.rank_all[sl]
        self._sl = sl
        self.state = torch.tensor([moco.index, 0], dtype=torch.int64, device=dev)
        self.qd = GF.queue_dtype_code(mem)
        self.ws = torch.zeros(GF.infonce_workspace_bytes(self.Bg, self.Ks, self.d, self.qd, self.algo),
                              dtype=torch.uint8, device=dev)
        self.graph = None
        self.launches_per_step = 0
        if link is not None:
            if (link.W, link.r, link.Bl, link.d) != (self.W, self.r, self.Bl, self.d):
                raise ValueError("PeerShardLink was built for another shape / group")
            # local-row outputs of the fused call (the NCCL path slices them out of the global-row buffers)
            self.lse, self.loss_rows = f32(self.Bl), f32(self.Bl)
            self.rank = torch.zeros(self.Bl, dtype=torch.int32, device=dev)
            self.hits = torch.zeros(2, dtype=torch.int32, device=dev)

    def _enqueue_work(self):
        import torch.distributed as dist
        m, sl = self.moco, self._sl
        dev = m.memory.device
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        if self.link is not None:
            x = self.link
            _lib.call("gca_shard_step_peer", ptr(self.inputs), ptr(m.memory), self.qd, self.Bl, m.K, self.d, 1.0 / m.T,
                      _lib.ALGO[self.algo], self.W, self.r, ptr(x.table), ptr(x.pstate), x.timeout_ms, ptr(self.state),
                      ptr(self.qk_all), ptr(self.loss), ptr(self.loss_rows), ptr(self.lse), ptr(self.pos), ptr(self.rank),
                      ptr(self.hits), ptr(self.dq), ptr(self.ws), self.ws.numel(), stream)
            return
        dist.all_gather_into_tensor(self.gathered, self.inputs, group=self.group)
        self.qk_all.view(2, self.W, self.Bl, self.d).copy_(self.gathered.transpose(0, 1))
        q_all, k_all = self.qk_all[0], self.qk_all[1]
        _lib.call("gca_infonce_shard_fwd", ptr(q_all), ptr(k_all), ptr(m.memory), self.qd, self.Bg, self.Ks, self.d,
                  1.0 / m.T, _lib.ALGO[self.algo], ptr(self.pos), ptr(self.stats[0]), ptr(self.stats[1]),
                  ptr(self.stats[2]), ptr(self.acc), ptr(self.ws), self.ws.numel(), stream)
        dist.all_gather_into_tensor(self.all_stats, self.stats, group=self.group)
        self.st.copy_(self.all_stats.transpose(0, 1))
        _lib.call("gca_infonce_shard_combine", ptr(self.st[0]), ptr(self.st[1]), ptr(self.st[2]), self.W, self.r, self.Bg,
                  self.d, ptr(self.pos), ptr(self.lse_all), ptr(self.loss_rows_all), ptr(self.rank_all), ptr(self.acc),
                  stream)
        dist.reduce_scatter_tensor(self.acc_loc, self.acc, group=self.group)
        _lib.call("gca_infonce_shard_finish", ptr(self.acc_loc), ptr(k_all[sl]), ptr(self.pos[sl]), ptr(self.lse_all[sl]),
                  ptr(self.loss_rows_all[sl]), self.Bl, self.d, 1.0 / m.T, ptr(self.dq), ptr(self.loss), stream)
        _lib.call("gca_enqueue_devptr", ptr(m.memory), self.qd, m.K, m.k_begin, m.k_begin + self.Ks, self.d, ptr(k_all),
                  self.Bg, ptr(self.state), stream)

    def capture(self):
        lib = _lib.load()
        m = self.moco
        dev = m.memory.device
        saved = m.memory.clone()                                    # the warm-up run below enqueues; undo it afterwards
        self.state.copy_(torch.tensor([m.index, 0], dtype=torch.int64))
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            self._enqueue_work()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        m.memory.copy_(saved)
        del saved
        self.state.copy_(torch.tensor([m.index, 0], dtype=torch.int64))
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        n0 = lib.gca_launch_count()
        with torch.cuda.graph(g):
            self._enqueue_work()
        self.launches_per_step = int(lib.gca_launch_count() - n0)
        self.graph = g
        return self

    def step(self, q=None, k=None):
        """Replay one step on the LOCAL q, k (copied into the static buffers when given)."""
        if self.graph is None:
            self.capture()
        if q is not None:
            self.q.copy_(q, non_blocking=True)
        if k is not None:
            self.k.copy_(k, non_blocking=True)
        self.graph.replay()
        self.moco.index = (self.moco.index + self.Bg) % self.moco.K
        return self.loss
