#!/usr/bin/env python
"""bench.py -- the reference's headline metric on B200: contrastive-head fwd+bwd steps/s at B=256, K=65536, d=128.

One step = what the trainer does around the head (tools/train_video_contrast_dis.py:411-428): MoCo logits against the
queue + InfoNCE loss + gradient w.r.t. q + top-1/top-5 of the positive + in-place enqueue of the gathered keys and
pointer update.  Arms:
  (default)          the B200 path: libgca_b200.so through gca_b200.GraphedMoCoStep (N=1) / ShardedRGBMoCo (N>1)
  --impl reference   the reference's CPU implementation of the same step (oracle port; the reference is pure Python
                     and does not exist on the GPU box), all host threads, rank 0 only
Prints ONE JSON line (contract in the task statement): value = device-timed steps/s with inputs resident in HBM,
e2e = same step from pinned HOST buffers with the H2D / D2H copies inside the timed region, roofline for the dominant
kernel (infonce_tc_kernel) timed alone, cpu_baseline = the oracle port timed on this box's host cores.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "video-graph-ssl_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "contrastive_head_fwd_bwd_steps_per_s"
UNIT = "steps/s (1 step = fwd+bwd+top-k+enqueue of 256 query rows against the 65536 x 128 queue)"
B, K, D, T = 256, 65536, 128, 0.07
POOL = 8                                  # distinct input batches cycled through the timed steps
L2_FLUSH_BYTES = 256 << 20                # > 126 MB L2


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained"),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(object):
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line): started before the warm-up so the first
    samples exist when the timed region begins; only samples whose timestamp falls inside [mark_start, mark_end] count."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.t0 = self.t1 = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def wait_first_sample(self, timeout=5.0):
        t = time.time()
        while self.p is not None and time.time() - t < timeout:
            if os.path.getsize(self.f.name) > 0:
                return
            time.sleep(0.02)

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.05)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [[c.strip() for c in r.split(",")] for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 9]
        os.unlink(self.f.name)
        if not rows:
            return out

        def ts(r):
            try:
                return datetime.datetime.strptime(r[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
            except ValueError:
                return None
        inside = [r for r in rows if self.t0 is not None and ts(r) is not None and self.t0 - 0.02 <= ts(r) <= self.t1 + 0.02]
        use = inside if inside else rows
        sm = sorted(float(r[2]) for r in use)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[6 + i] == "Active" for r in use)]
        out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=float(use[0][3]), reasons=reasons, samples=len(use),
                   samples_in_timed_region=len(inside), power_w_max=max(float(r[4]) for r in use))
        return out


def synthetic_batches(gen_seed, n, rows_q, rows_k, device=None, pin=False):
    """POOL batches of L2-normalised rows, generated on the CPU with a fixed seed."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(gen_seed)
    out = []
    for _ in range(n):
        packed = torch.cat([F.normalize(torch.randn(rows_q, D, generator=g)), F.normalize(torch.randn(rows_q, D, generator=g)),
                            F.normalize(torch.randn(rows_k, D, generator=g))])
        if pin:
            packed = packed.pin_memory()
        out.append(packed if device is None else packed.to(device))
    return out


# ----------------------------------------------------------------------------------------------- CPU (reference arm)
def cpu_head_rate(steps, warmup, budget_s):
    """The oracle port of the reference step on the host cores: fp32, all threads, same shapes."""
    import torch
    import torch.nn.functional as F
    import oracle
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(1)
    mem = F.normalize(torch.randn(K, D, generator=g))
    batches = synthetic_batches(2, 4, B, B)
    idx, times = 0, []
    t_start = time.perf_counter()
    for i in range(warmup + steps):
        pk = batches[i % len(batches)]
        q = pk[:B].clone().requires_grad_(True)
        t0 = time.perf_counter()
        _, _, idx, _ = oracle.infonce.reference_head_step(q, pk[B:2 * B], mem, idx, T, all_k=pk[2 * B:])
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if i >= warmup + 2 and time.perf_counter() - t_start > budget_s:
            break
    mean = sum(times) / len(times)
    return 1.0 / mean, mean * 1e3, len(times), torch.get_num_threads()


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference(args, rank, world):
    if rank != 0:
        return
    steps = min(args.steps, 400)
    rate, ms, n, cores = cpu_head_rate(steps, min(args.warmup, 5), budget_s=150.0)
    sample = "%d full-size steps (B=%d, K=%d, d=%d, fp32) of the oracle port of the reference step, %s" % (n, B, K, D, cpu_model())
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": n, "warmup": min(args.warmup, 5),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "moco_head_B256_K65536_d128", "B": B, "K": K, "d": D, "T": T, "device": "cpu"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- B200 arm, N = 1
def run_single(args):
    import torch
    import torch.nn.functional as F
    import gca_b200
    from gca_b200 import _lib
    from gca_b200.graphed import GraphedMoCoStep
    from gca_b200 import functional as GF
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.manual_seed(1)
    moco = gca_b200.RGBMoCo(D, K=K, T=T, queue_dtype="bf16").to(dev)
    batches = synthetic_batches(2, POOL, B, B, device=dev)
    state = torch.tensor([0, 0], dtype=torch.int64, device=dev)
    steps_g = []
    for i in range(POOL):                                  # one captured step per pool entry, shared queue + ring pointer
        s = GraphedMoCoStep(moco, B, B, state=state)
        s.inputs.copy_(batches[i])
        steps_g.append(s)
    for s in steps_g:
        s.capture()
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]

    def one(i):
        steps_g[i % POOL].step()

    sampler = ClockSampler(0)
    for i in range(args.warmup):
        flush.fill_(i & 1)
        one(i)
    sampler.wait_first_sample()
    torch.cuda.synchronize()
    sampler.mark_start()
    t0 = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(i & 1)                                 # L2 flush between timed iterations (not inside the events)
        ev[i][0].record()
        one(args.warmup + i)
        ev[i][1].record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    sampler.mark_end()
    clocks = sampler.stop()
    dev_ms = sorted(a.elapsed_time(b) for a, b in ev)
    ms_per_step = sum(dev_ms) / len(dev_ms)
    launches = steps_g[0].launches_per_step * args.steps
    loss_last = float(steps_g[(args.warmup + args.steps - 1) % POOL].loss)

    # ---- end to end: pinned host inputs -> H2D -> step -> D2H of loss/top-k/dq, synchronised every step
    # (each pool entry owns a pinned input batch and a pinned result buffer; the two copies are nodes of the step's graph, so
    # a step is one graph launch + one stream synchronise, like `loss.item()` in the reference loop)
    host_in = synthetic_batches(3, POOL, B, B, pin=True)
    g0 = steps_g[0]
    host_outs = [torch.empty_like(g0.outputs, device="cpu").pin_memory() for _ in range(POOL)]
    zc_in = os.environ.get("GCA_BENCH_ZC_IN", "1") == "1"  # first kernel reads q|k (and the enqueue CTAs all_k) from pinned host memory
    for i in range(POOL):
        steps_g[i].capture_host_io(host_in[i], host_outs[i], zero_copy_out=True, zero_copy_in=zc_in)
    e2e_t = []
    e2e_steps = min(args.steps, 2000)
    e2e_loss = 0.0
    for i in range(args.warmup + e2e_steps):
        flush.fill_(i & 1)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        steps_g[i % POOL].step_host_io()
        torch.cuda.synchronize()
        e2e_loss = float(host_outs[i % POOL][0])           # the step's result, read on the host
        if i >= args.warmup:
            e2e_t.append(time.perf_counter() - t1)
    e2e_ms = sum(e2e_t) / len(e2e_t) * 1e3
    h2d = g0.inputs.numel() * 4
    d2h = g0.outputs.numel() * 4

    # ---- dominant kernel alone (roofline): the queue-streaming tcgen05 kernel, cold L2
    q, k = batches[0][:B].contiguous(), batches[0][B:2 * B].contiguous()
    ws = GF.workspace(dev, GF.infonce_workspace_bytes(B, K, D, 1, "tcgen05"), "bench")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    # first call prepares the bf16 q block / positive logits in `ws`; the timed launches (flag bit 1) run the streaming kernel
    # alone.  The launch is captured in a CUDA graph (the product step launches it that way too); cold L2 before every launch.
    _lib.call("gca_infonce_partials", _lib.ptr(q), _lib.ptr(k), _lib.ptr(moco.memory), 1, B, K, D, 1.0 / T, 2, 1,
              _lib.ptr(ws), ws.numel(), st)
    torch.cuda.synchronize()
    # the two timing events are recorded INSIDE the captured graph (external event-record nodes), directly around the kernel
    # node, so the interval is the kernel's own duration; a replay-level event pair would add ~6 us of launch latency
    kgraph = torch.cuda.CUDAGraph()
    try:
        ka, kb = torch.cuda.Event(enable_timing=True, external=True), torch.cuda.Event(enable_timing=True, external=True)
        in_graph_events = True
    except TypeError:
        ka = kb = None
        in_graph_events = False
    with torch.cuda.graph(kgraph):
        if in_graph_events:
            ka.record()
        _lib.call("gca_infonce_partials", _lib.ptr(q), _lib.ptr(k), _lib.ptr(moco.memory), 1, B, K, D, 1.0 / T, 2, 3,
                  _lib.ptr(ws), ws.numel(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        if in_graph_events:
            kb.record()
    kt = []
    for i in range(110):
        flush.fill_(i & 1)
        if in_graph_events:
            kgraph.replay()
            torch.cuda.synchronize()
            t_k = ka.elapsed_time(kb)
        else:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            kgraph.replay()
            b.record()
            torch.cuda.synchronize()
            t_k = a.elapsed_time(b)
        if i >= 10:
            kt.append(t_k)
    k_ms_single = sum(kt) / len(kt)
    # Average launch duration in a train of launches: NQ back-to-back launches, each over its OWN copy of the queue (NQ x 16.8 MB
    # = 201 MB > the 126 MB L2, so every launch streams its queue from HBM), bracketed by one in-graph event pair.  This
    # takes the ~6 us that an event pair adds around a single 14 us kernel out of the figure; the single-launch number is
    # reported next to it.
    k_ms = k_ms_single
    train = None
    if in_graph_events:
        NQ = 12
        queues = [moco.memory.clone() for _ in range(NQ)]
        tgraph = torch.cuda.CUDAGraph()
        ta, tb = torch.cuda.Event(enable_timing=True, external=True), torch.cuda.Event(enable_timing=True, external=True)
        with torch.cuda.graph(tgraph):
            ta.record()
            for qu in queues:
                _lib.call("gca_infonce_partials", _lib.ptr(q), _lib.ptr(k), _lib.ptr(qu), 1, B, K, D, 1.0 / T, 2, 3,
                          _lib.ptr(ws), ws.numel(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
            tb.record()
        tt = []
        for i in range(35):
            flush.fill_(i & 1)
            tgraph.replay()
            torch.cuda.synchronize()
            if i >= 5:
                tt.append(ta.elapsed_time(tb) / NQ)
        k_ms = sum(tt) / len(tt)
        train = {"launches_per_train": NQ, "distinct_queue_bytes": NQ * K * D * 2, "trains": len(tt)}
        del queues
    pk = peaks()
    flops = 4.0 * B * K * D                                 # single pass: S = q Q^T and O += P Q
    alg_bytes = K * D * 2 + 2 * B * D * 4 + 74 * B * (D + 3) * 4   # queue once + q,k + split partials written
    t_tensor = flops / (pk["bf16_tflops"] * 1e12)
    t_hbm = alg_bytes / (pk["hbm_gbs"] * 1e9)
    ach_tf = flops / (k_ms * 1e-3) / 1e12
    roof = {"bound": "tensor" if t_tensor >= t_hbm else "hbm", "achieved": ach_tf, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
            "frac": ach_tf / pk["bf16_tflops"], "traffic": None, "kernel": "infonce_tc_kernel<acc,online-max>", "kernel_ms": k_ms,
            "alg_flops": flops, "alg_bytes": alg_bytes, "hbm_achieved_gbs": alg_bytes / (k_ms * 1e-3) / 1e9,
            "hbm_frac": alg_bytes / (k_ms * 1e-3) / 1e9 / pk["hbm_gbs"], "peaks": pk["source"] + ", burst bf16 (kernel timed alone)",
            "timing": ("in-graph CUDA events around a train of %d launches over %d distinct queue copies (%.0f MB > L2), time / launches"
                       % (train["launches_per_train"], train["launches_per_train"], train["distinct_queue_bytes"] / 1e6))
                      if train else "CUDA events around a graph replay",
            "kernel_ms_single_launch": k_ms_single, "frac_single_launch": flops / (k_ms_single * 1e-3) / 1e12 / pk["bf16_tflops"],
            "l2": "flushed before every train; every launch of a train reads a queue copy that is not L2-resident"}
    prof = os.path.join(ROOT, "profiles", "r01_dram_traffic.json")
    if os.path.exists(prof):
        try:
            roof["traffic"] = json.load(open(prof)).get("infonce_tc_kernel_dram_bytes")
        except (ValueError, OSError):
            pass

    cpu_rate, cpu_ms, cpu_n, cores = cpu_head_rate(60, 2, budget_s=15.0) if not args.no_cpu else (None, None, 0, 0)
    line = {
        "metric": METRIC, "value": 1e3 / ms_per_step, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": "moco_head_B256_K65536_d128", "B": B, "K": K, "d": D, "T": T, "queue_dtype": "bf16",
                   "algo": "tcgen05 single-pass (loss + dq in one queue sweep)", "cuda_graph": True, "input_pool": POOL,
                   "l2": "flushed between timed iterations (256 MiB write, outside the per-step events)",
                   "timing": "per-step CUDA events on the launch stream; value = 1000 / mean(ms)"},
        "ms_per_step_median": dev_ms[len(dev_ms) // 2], "wall_s_total": wall, "loss_last": loss_last,
        "clocks": clocks,
        "e2e": {"value": 1e3 / e2e_ms, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms,
                "path": ("GraphedMoCoStep.step_host_io(): the step's first kernel reads q|k from the pinned host buffer over PCIe "
                         "(zero-copy, each once), its enqueue CTAs read all_k from it, " if zc_in else
                         "GraphedMoCoStep.step_host_io(): pinned host q|k|all_k -> H2D copy -> step (C ABI), ") +
                        "its last kernel stores loss|top-k hits|dq straight into the pinned host result buffer -- one graph "
                        "launch, stream synchronised and the loss read on the host every step",
                "zero_copy_in": zc_in,
                "loss_last": e2e_loss},
        "gpu_launches": launches,
        "roofline": roof,
        "cpu_baseline": None if args.no_cpu else {
            "value": cpu_rate, "unit": UNIT, "cores": cores, "kind": "port", "ms_per_step": cpu_ms,
            "sample": "%d full-size steps (B=%d, K=%d, d=%d, fp32) of the oracle port of the reference step, %s" % (cpu_n, B, K, D, cpu_model())},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- B200 arm, N > 1
def run_multi(args, rank, world, local_rank):
    """Weak scaling, the reference's own data parallelism: every GPU scores its 256 rows against a replicated 65536-row
    queue; the keys of all ranks are all-gathered (NCCL over NVLink, overlapped with the queue sweep) and enqueued on every
    replica.  value = n_gpus x steps/s.  The K-sharded variant (queue 2^20 x 128 split over the ranks, BASELINE config 4) is
    timed in the same run and reported under "sharded_k1m"."""
    import torch
    import torch.distributed as dist
    import gca_b200
    from gca_b200 import _lib
    from gca_b200.graphed import GraphedReplicaStep
    lib = _lib.load()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1)                                     # identical queue on every replica (upstream: broadcast from rank 0)
    moco = gca_b200.RGBMoCo(D, K=K, T=T, queue_dtype="bf16").to(dev)
    dist.broadcast(moco.memory, 0)
    pool = 4
    batches = synthetic_batches(100 + rank, pool, B, 0, device=dev)
    state = torch.tensor([0, 0], dtype=torch.int64, device=dev)
    # key exchange: one peer-memory kernel per step (gca_keys_exchange over NVLink); NCCL all-gather if symmetric memory
    # cannot be set up on this box (GCA_BENCH_EXCHANGE=nccl forces it).  All ranks agree on the mode.
    exchange, why = None, "forced by GCA_BENCH_EXCHANGE"
    xmode = os.environ.get("GCA_BENCH_EXCHANGE", "fused")          # fused | p2p (stand-alone kernel) | nccl
    if xmode in ("fused", "p2p"):
        try:
            from gca_b200.peer import PeerKeyExchange
            exchange = PeerKeyExchange(B, D, device=dev, timeout_ms=30000)
            probe = batches[0][B:2 * B].contiguous()
            got = torch.empty(world * B, D, device=dev)
            want = torch.empty(world * B, D, device=dev)
            exchange(probe, got)
            dist.all_gather_into_tensor(want, probe)
            torch.cuda.synchronize()
            exchange.check()
            if not torch.equal(got, want):
                raise RuntimeError("peer exchange disagrees with NCCL all-gather")
            okf = 1
        except Exception as e:
            okf, why = 0, "%s: %s" % (type(e).__name__, e)
        agree = torch.tensor([okf], device=dev)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN)
        if int(agree) == 0:
            if rank == 0:
                print("peer-memory key exchange unavailable (%s); using NCCL all-gather" % why, file=sys.stderr)
            exchange = None
    steps_g = []
    for i in range(pool):
        s = GraphedReplicaStep(moco, B, state=state, exchange=exchange, fuse_exchange=(xmode == "fused"))
        s.inputs[:2 * B].copy_(batches[i])
        steps_g.append(s)
    graphed = True
    torch.cuda.synchronize()
    dist.barrier()                                           # ranks enter the (peer-synchronised) warm-up launches together
    try:
        for s in steps_g:
            s.capture()
    except Exception as e:                                   # NCCL capture unavailable: run the same work eagerly
        graphed = False
        if rank == 0:
            print("CUDA-graph capture of the NCCL step failed (%s); running eagerly" % e, file=sys.stderr)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def one(i):
        s = steps_g[i % pool]
        if graphed:
            s.step()
        else:
            s._enqueue_work(st)
            moco.index = (moco.index + s.N) % K

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for i in range(args.warmup):
        flush.fill_(i & 1)
        one(i)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    if sampler:
        sampler.wait_first_sample()
    n0 = lib.gca_launch_count()
    torch.cuda.synchronize()
    dist.barrier()
    if sampler:
        sampler.mark_start()
    t0 = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(i & 1)
        ev[i][0].record()
        one(args.warmup + i)
        ev[i][1].record()
    torch.cuda.synchronize()
    dist.barrier()
    wall = time.perf_counter() - t0
    if sampler:
        sampler.mark_end()
    clocks = sampler.stop() if sampler else None
    launches = (steps_g[0].launches_per_step * args.steps) if graphed else int(lib.gca_launch_count() - n0)
    tot = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], device=dev)
    dist.all_reduce(tot, op=dist.ReduceOp.MAX)              # max over ranks of the device-timed total
    ms_per_step = float(tot) / args.steps
    loss_last = float(steps_g[(args.warmup + args.steps - 1) % pool].loss)
    # replicas must stay bit-identical: compare a checksum of the queue and the ring pointer across ranks
    chk = torch.stack([moco.memory.float().sum().double(), state[0].double()])
    chk_all = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(chk_all, chk)
    consistent = all(torch.equal(c, chk_all[0]) for c in chk_all)

    # end to end: pinned host q|k per rank -> H2D -> step -> D2H loss|hits|dq, synchronised per step
    host_in = synthetic_batches(200 + rank, pool, B, 0, pin=True)
    g0 = steps_g[0]
    host_outs = [torch.empty_like(g0.outputs, device="cpu").pin_memory() for _ in range(pool)]
    host_io = graphed
    if graphed:
        try:                                                 # same path as at one GPU: no copy nodes, one graph launch per step
            torch.cuda.synchronize()
            dist.barrier()
            for i in range(pool):
                steps_g[i].capture_host_io(host_in[i], host_outs[i], zero_copy_out=True,
                                           zero_copy_in=(exchange is not None and xmode == "fused"))   # (NCCL gathers device rows)
        except Exception as e:
            host_io = False
            if rank == 0:
                print("host-io graph unavailable (%s); copying around the step" % e, file=sys.stderr)
    e2e_t = []
    e2e_steps = min(args.steps, 1000)
    for i in range(args.warmup + e2e_steps):
        flush.fill_(i & 1)
        torch.cuda.synchronize()
        dist.barrier()
        t1 = time.perf_counter()
        if host_io:
            steps_g[i % pool].step_host_io()
        else:
            g0.inputs[:2 * B].copy_(host_in[i % pool], non_blocking=True)
            if graphed:
                g0.step()
            else:
                g0._enqueue_work(st)
            host_outs[0].copy_(g0.outputs, non_blocking=True)
        torch.cuda.synchronize()
        if i >= args.warmup:
            e2e_t.append(time.perf_counter() - t1)
    e2e = torch.tensor([sum(e2e_t) / len(e2e_t) * 1e3], device=dev)
    dist.all_reduce(e2e, op=dist.ReduceOp.MAX)

    if exchange is not None:
        # an in-kernel wait for a peer that timed out (a rank stalled for > 30 s) invalidates the run: every rank agrees on it
        # and the whole measurement is repeated once with the NCCL all-gather instead
        bad = torch.tensor([1 if int(exchange.xstate[2]) != 0 else 0], device=dev)
        dist.all_reduce(bad, op=dist.ReduceOp.MAX)
        if int(bad) != 0:
            if rank == 0:
                print("peer-memory key exchange timed out waiting for a rank; repeating the run with NCCL", file=sys.stderr)
            dist.barrier()
            torch.cuda.synchronize()
            os.environ["GCA_BENCH_EXCHANGE"] = "nccl"
            sys.stdout.flush()
            sys.stderr.flush()
            os.execv(sys.executable, [sys.executable] + sys.argv)
    sharded = time_sharded_k1m(rank, world, dev, flush) if not args.no_sharded else None
    if rank == 0:
        line = {
            "metric": METRIC, "value": world * 1e3 / ms_per_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "moco_head_B256perGPU_K65536_d128_replicas", "B_per_gpu": B, "B_global": B * world, "K": K, "d": D,
                       "T": T, "queue_dtype": "bf16", "enqueued_rows_per_step": B * world,
                       "parallelism": "dp%d: replicated queue (the reference's scheme); key all-gather (%s) overlapped with "
                                      "the queue sweep; every replica enqueues all %d keys"
                                      % (world, "NCCL" if exchange is None else
                                         "NVLink peer-memory stores fused into the step's own launches" if xmode == "fused" else
                                         "gca_keys_exchange: one NVLink peer-memory kernel per step", B * world),
                       "key_exchange": "nccl" if exchange is None else ("fused_p2p" if xmode == "fused" else "p2p_kernel"),
                       "cuda_graph": graphed, "l2": "flushed between timed iterations (256 MiB write, outside the per-step events)",
                       "timing": "per-step CUDA events, total = max over ranks; value = n_gpus * 1000 / ms_per_step "
                                 "(each global step processes n_gpus x 256 rows)"},
            "wall_s_total": wall, "loss_last": loss_last, "replicas_consistent": consistent, "clocks": clocks,
            "e2e": {"value": world * 1e3 / float(e2e), "unit": UNIT, "h2d_bytes_per_step": 2 * B * D * 4,
                    "d2h_bytes_per_step": g0.outputs.numel() * 4, "ms_per_step": float(e2e),
                    "path": "GraphedReplicaStep.step_host_io(): one graph launch per step, q|k read from / results stored to pinned "
                            "host buffers by the step's own kernels, host barrier + stream sync every step" if host_io else
                            "H2D copy -> step -> D2H copy, host barrier + stream sync every step"},
            "gpu_launches": launches,
            "sharded_k1m": sharded,
        }
        print(json.dumps(line), flush=True)
    # Leave without tearing the communicator down: destroy_process_group() after NCCL work was captured into CUDA graphs
    # has been seen to block forever (round 1, 2 GPUs); every rank is past its last collective here.
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def time_sharded_k1m(rank, world, dev, flush, steps=200, warmup=10):
    """BASELINE config 4: queue 2^20 x 128 (bf16) split along K over the ranks, 256 rows per GPU; per-shard online-softmax
    partials merged with NCCL.  One CUDA graph per step (gca_b200.graphed.GraphedShardedStep: 3 collectives + kernels);
    the first step is cross-checked against the eager ShardedRGBMoCo path.  Per-step device time with an L2 flush
    before every step, max over ranks."""
    import torch
    import torch.distributed as dist
    import gca_b200
    from gca_b200.dist import ShardedRGBMoCo
    from gca_b200.graphed import GraphedShardedStep
    K1 = 1 << 20
    torch.manual_seed(1)
    moco = ShardedRGBMoCo(D, K=K1, T=T, queue_dtype="bf16", device=dev)
    crit = gca_b200.NCESoftmaxLoss()
    batches = synthetic_batches(300 + rank, 2, B, 0, device=dev)
    # eager reference step on a copy of the shard
    shard0 = moco.memory.clone()
    q = batches[0][:B].clone().requires_grad_(True)
    out, _ = moco(q, batches[0][B:2 * B])
    loss_eager = crit(out)
    loss_eager.backward()
    torch.cuda.synchronize()
    mem_eager = moco.memory.clone()
    moco.memory.copy_(shard0)
    moco.index = 0
    del shard0
    gs = GraphedShardedStep(moco, B).capture()
    gs.step(batches[0][:B], batches[0][B:2 * B])
    torch.cuda.synchronize()
    same = bool(torch.equal(gs.loss, loss_eager.detach().reshape(1)) and torch.equal(gs.dq, q.grad)
                and torch.equal(moco.memory, mem_eager))
    del mem_eager
    ev = []
    for i in range(warmup + steps):
        flush.fill_(i & 1)
        pk = batches[i % 2]
        gs.q.copy_(pk[:B])
        gs.k.copy_(pk[B:2 * B])
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        gs.step()
        b.record()
        if i >= warmup:
            ev.append((a, b))
    torch.cuda.synchronize()
    tot = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], device=dev)
    dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    ok = torch.tensor([1 if same else 0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    ms = float(tot) / steps
    flops = 4.0 * (B * world) * (K1 / world) * D
    return {"K": K1, "rows_per_gpu": B, "rows_global": B * world, "shard_rows": K1 // world, "ms_per_step": ms,
            "global_rows_per_s": B * world / ms * 1e3, "per_gpu_tflops": flops / (ms * 1e-3) / 1e12, "cuda_graph": True,
            "launches_per_step": gs.launches_per_step, "collectives_per_step": 3,
            "graph_equals_eager_path": bool(int(ok))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the extra K-sharded 2^20-row measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if world > 1:
        return run_multi(args, rank, world, local_rank)
    if args.gpus > 1:
        print("bench.py --gpus %d must be launched with torch.distributed.run (one rank per GPU)" % args.gpus, file=sys.stderr)
        sys.exit(2)
    return run_single(args)


if __name__ == "__main__":
    main()
