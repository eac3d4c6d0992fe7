#!/usr/bin/env python
"""bench.py -- the reference's headline metric on B200: contrastive-head fwd+bwd steps/s at B=256, K=65536, d=128.

One step = what the trainer does around the head (tools/train_video_contrast_dis.py:411-428): MoCo logits against the
queue + InfoNCE loss + gradient w.r.t. q + top-1/top-5 of the positive + in-place enqueue of the gathered keys and
pointer update.  Arms:
  (default)          the B200 path: libgca_b200.so through gca_b200.GraphedMoCoStep (N=1) / ShardedRGBMoCo (N>1)
  --impl reference   the reference's CPU implementation of the same step (oracle port; the reference is pure Python
                     and does not exist on the GPU box), all host threads, rank 0 only
Prints ONE JSON line (contract in the task statement): value = device-timed steps/s with inputs resident in HBM -- EXACTLY
`--steps` steps back to back between one pair of CUDA events, every step on its own replica of the queue so that the
working set (318 MB) exceeds the L2 -- e2e = same step from pinned HOST buffers with the H2D / D2H traffic inside the
timed region, roofline for the dominant kernel (infonce_tcx_kernel) timed alone, cpu_baseline = the oracle port timed on
this box's host cores.  `ms_per_step_isolated` keeps round 1's figure next to it: one step between its own event pair
after an explicit L2 flush (it additionally contains ~2.5 us of graph-launch latency that back-to-back steps overlap).
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
QPOOL = 12                                # distinct queue replicas cycled through the timed steps: 12 x (16.8 MB queue + 9.7 MB
                                          # split partials) = 318 MB > the 126 MB L2, so every step streams its queue from HBM
L2_FLUSH_BYTES = 256 << 20                # > 126 MB L2 (isolated-step and e2e legs)


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


# ----------------------------------------------------------------------------------------------- secondary figures
def _time_graph(fn, flush, iters=20, warm=3):
    """fn() captured once in a CUDA graph, replayed with an L2 flush before every replay; mean device ms per replay."""
    import torch
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    ts = []
    for i in range(warm + iters):
        flush.fill_(i & 1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(a.elapsed_time(b))
    return sum(ts) / len(ts)


def secondary_measurements(dev, flush, pk):
    """Driver-visible figures for the other SURVEY 8 configs, each a few replays (cold L2) of the C-ABI call: the fp32 parity
    mode of the head, the head at K = 2^20 on one GPU (the N = 1 point of the K-sharded series), the temporal-graph core
    forward + backward at BASELINE config 3 with its fraction of the measured HBM peak, SimSiam's D, retrieval (config 5)."""
    import torch
    import torch.nn.functional as F
    from gca_b200 import functional as GF
    out = {}
    g = torch.Generator().manual_seed(5)
    unit = lambda n: F.normalize(torch.randn(n, D, generator=g)).to(dev)
    q, k = unit(B), unit(B)
    try:
        mem32 = unit(K)
        ms = _time_graph(lambda: GF.infonce_forward(q, k, mem32, T, algo="tc32", want_grad=True), flush, iters=12, warm=3)
        ms_ffma = _time_graph(lambda: GF.infonce_forward(q, k, mem32, T, algo="ffma", want_grad=True), flush, iters=6, warm=2)
        out["fp32_mode"] = {"what": "fused head, fp32 queue, fp32-grade logits on tcgen05 (GCA_ALGO_TC32: exact 3-way bf16 split, 6 piece "
                                    "products per logit; 1e-5 parity mode), B=256 K=65536; includes the per-call split of the queue",
                            "ms": ms, "steps_per_s": 1e3 / ms, "tflops_bf16_mma": (6 * 2 + 2) * B * K * D / (ms * 1e-3) / 1e12,
                            "cuda_core_ffma_ms": ms_ffma}
        del mem32
    except Exception as e:                                   # a secondary figure must never take the headline down
        out["fp32_mode"] = {"error": "%s: %s" % (type(e).__name__, e)}
    try:
        K1 = 1 << 20
        mem1 = F.normalize(torch.randn(K1, D, generator=g)).to(torch.bfloat16).to(dev)
        ms = _time_graph(lambda: GF.infonce_forward(q, k, mem1, T, algo="tcgen05", want_grad=True), flush, iters=10, warm=2)
        fl = 4.0 * B * K1 * D
        out["head_k1m_1gpu"] = {"what": "fused head, bf16 queue 2^20 x 128 on ONE GPU, B=256 (N=1 point of the K-sharded series)",
                                "ms": ms, "global_rows_per_s": B / ms * 1e3, "tflops": fl / (ms * 1e-3) / 1e12,
                                "frac_of_bf16_peak": fl / (ms * 1e-3) / 1e12 / pk["bf16_tflops"]}
        del mem1
    except Exception as e:
        out["head_k1m_1gpu"] = {"error": "%s: %s" % (type(e).__name__, e)}
    try:
        Bv, C, Tt, HW, Cq, S = 128, 192, 8, 196, 96, 49      # S3D base.5 map [128,192,8,14,14], sub-sampled projections (7x7)
        gq = torch.randn(Bv, Cq, Tt, S, device=dev) * 0.02
        gk = torch.randn(Bv, Cq, Tt, S, device=dev) * 0.02
        sup = torch.randn(Bv, C, Tt, HW, device=dev)
        u = torch.rand(Bv, Tt, Tt, device=dev)
        dy = torch.randn(Bv, C, Tt, HW, device=dev)
        gq.requires_grad_(True); gk.requires_grad_(True); sup.requires_grad_(True)

        def fb():
            y = GF.graph_core(gq, gk, sup, u)[0]
            torch.autograd.grad(y, (gq, gk, sup), dy)
        ms = _time_graph(fb, flush, iters=10, warm=2)
        per_dir = 4.0 * (2 * Tt * Cq * S + 2 * C * Tt * HW + 2 * Tt * Tt) * Bv          # SURVEY 8d, per direction
        bwd = 4.0 * (2 * Tt * Cq * S * 2 + 3 * C * Tt * HW + 4 * Tt * Tt) * Bv          # reads gq,gk,sup,dy,T^2 terms; writes d_gq,d_gk,d_sup
        out["graph_head_c3"] = {"what": "temporal-graph core fwd+bwd (gca_graph_fwd + gca_graph_bwd), 128 x [192,8,14,14]", "ms": ms,
                                "videos_per_s": Bv / ms * 1e3, "alg_bytes": per_dir + bwd,
                                "hbm_frac": (per_dir + bwd) / (ms * 1e-3) / 1e9 / pk["hbm_gbs"]}
    except Exception as e:
        out["graph_head_c3"] = {"error": "%s: %s" % (type(e).__name__, e)}
    try:
        rng = torch.Generator().manual_seed(0)
        gal, qry = torch.randn(13320, 512, generator=rng).to(dev), torch.randn(3783, 512, generator=rng).to(dev)
        ms = _time_graph(lambda: GF.cosine_topk(qry, gal, 50, normalize=True), flush, iters=8, warm=2)
        out["retrieval_c5"] = {"what": "cosine top-50, 3783 x 13320 x 512 (gca_sim_topk)", "ms": ms,
                               "tflops_bf16_equiv": 6 * 2.0 * 3783 * 13320 * 512 / (ms * 1e-3) / 1e12}
    except Exception as e:
        out["retrieval_c5"] = {"error": "%s: %s" % (type(e).__name__, e)}
    try:
        # instance-bank mode (MEM_TYPE 'bank', lib/memory/mem_bank.py): bsz 256, 16384 sampled negatives, 240k-clip bank
        Bb, Kb, nb = 256, 16384, 240000
        bank = F.normalize(torch.randn(nb, D, device=dev))
        xb = F.normalize(torch.randn(Bb, D, device=dev)).requires_grad_(True)
        yb = torch.randperm(nb, device=dev)[:Bb]
        ib = torch.randint(0, nb, (Bb, Kb + 1), device=dev)
        ib[:, 0] = yb
        wb = torch.randn(Bb, Kb + 1, device=dev)

        def bank_step():
            lg = GF.bank_logits(xb, bank, ib, T)
            torch.autograd.grad(lg, (xb,), wb)
            GF.bank_update_(bank, xb, yb, 0.5)
        ms = _time_graph(bank_step, flush, iters=8, warm=2)
        gathered = 2.0 * Bb * (Kb + 1) * D * 4
        out["instance_bank"] = {"what": "bank logits + dx + momentum update (gca_bank_*), bsz 256, K 16384, d 128, n_data 240000",
                                "ms": ms, "steps_per_s": 1e3 / ms, "alg_bytes": gathered,
                                "hbm_frac": gathered / (ms * 1e-3) / 1e9 / pk["hbm_gbs"]}
    except Exception as e:
        out["instance_bank"] = {"error": "%s: %s" % (type(e).__name__, e)}
    return out


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
    # the launching thread and the pinned host buffers of the e2e leg go to the GPU's own NUMA node (gca_b200/affinity.py);
    # undone before the CPU baseline leg, which uses every host core
    from gca_b200 import affinity
    aff = affinity.bind_cpu_to_device(0) if os.environ.get("GCA_BENCH_NO_BIND") != "1" else {"bound": False, "source": "off"}
    torch.manual_seed(1)
    batches = synthetic_batches(2, POOL, B, B, device=dev)
    # QPOOL replicas of the queue, one captured step each (the step's inputs cycle through POOL batches): consecutive timed
    # steps touch disjoint 26.5 MB working sets, 318 MB in all, so no step finds its queue in the 126 MB L2
    mocos = [gca_b200.RGBMoCo(D, K=K, T=T, queue_dtype="bf16").to(dev) for _ in range(QPOOL)]
    moco = mocos[0]
    steps_g = []
    for i in range(QPOOL):
        # want_rank=False: the step reports what the trainer logs -- the top-1 / top-5 hit counts (train...:428) -- not 256 ranks
        sg = GraphedMoCoStep(mocos[i], B, B, want_rank=False)
        sg.inputs.copy_(batches[i % POOL])
        steps_g.append(sg)
    for sg in steps_g:
        sg.capture()
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    def one(i):
        steps_g[i % QPOOL].step()

    sampler = ClockSampler(0)
    for i in range(args.warmup):
        one(i)
    sampler.wait_first_sample()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.mark_start()
    t0 = time.perf_counter()
    ev0.record()
    for i in range(args.steps):                            # EXACTLY K steps, back to back, between one event pair
        one(args.warmup + i)
    ev1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    sampler.mark_end()
    clocks = sampler.stop()
    ms_per_step = ev0.elapsed_time(ev1) / args.steps
    launches = (steps_g[0].plan.launches if steps_g[0].plan is not None else steps_g[0].launches_per_step) * args.steps
    loss_last = float(steps_g[(args.warmup + args.steps - 1) % QPOOL].loss)
    # round 1's figure for continuity: one step between its own event pair, L2 flushed before it
    iso = []
    for i in range(60):
        flush.fill_(i & 1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        one(i)
        b.record()
        torch.cuda.synchronize()
        if i >= 10:
            iso.append(a.elapsed_time(b))
    ms_isolated = sum(iso) / len(iso)

    # ---- end to end: pinned host inputs -> H2D -> step -> D2H of loss/top-k/dq, synchronised every step
    # (each pool entry owns a pinned input batch and a pinned result buffer; the two copies are nodes of the step's graph, so
    # a step is one graph launch + one stream synchronise, like `loss.item()` in the reference loop)
    host_in = synthetic_batches(3, POOL, B, B, pin=True)
    g0 = steps_g[0]
    host_outs = [torch.empty_like(g0.outputs, device="cpu").pin_memory() for _ in range(POOL)]
    assert POOL <= QPOOL
    zc_in = os.environ.get("GCA_BENCH_ZC_IN", "1") == "1"  # first kernel reads q|k (and the enqueue CTAs all_k) from pinned host memory
    for i in range(POOL):
        steps_g[i].capture_host_io(host_in[i], host_outs[i], zero_copy_out=True, zero_copy_in=zc_in)
    e2e_t = []
    e2e_steps = min(args.steps, 2000)
    e2e_loss = 0.0
    e2e_sync = os.environ.get("GCA_BENCH_E2E_SYNC") == "1"     # A/B: cudaDeviceSynchronize instead of the completion word
    for i in range(args.warmup + e2e_steps):
        flush.fill_(i & 1)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        if e2e_sync:
            steps_g[i % POOL].step_host_io()
            torch.cuda.synchronize()
        else:
            steps_g[i % POOL].step_host_io(wait=True)      # returns when the results are in the host buffer (completion word)
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
    # ALGORITHMIC bytes of the single-pass formulation: queue once + q, k, dq + enqueued rows + per-row scalars
    alg_bytes = K * D * 2 + 12 * B * D + B * D * (4 + 2) + 16 * B
    t_tensor = flops / (pk["bf16_tflops"] * 1e12)
    t_hbm = alg_bytes / (pk["hbm_gbs"] * 1e9)
    ach_tf = flops / (k_ms * 1e-3) / 1e12
    roof = {"bound": "tensor" if t_tensor >= t_hbm else "hbm", "achieved": ach_tf, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
            "frac": ach_tf / pk["bf16_tflops"], "traffic": None, "kernel": "infonce_tcx_kernel", "kernel_ms": k_ms,
            "alg_flops": flops, "alg_bytes": alg_bytes, "hbm_achieved_gbs": alg_bytes / (k_ms * 1e-3) / 1e9,
            "hbm_frac": alg_bytes / (k_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
            "step_frac": flops / (ms_per_step * 1e-3) / 1e12 / pk["bf16_tflops"],
            "step_hbm_frac": alg_bytes / (ms_per_step * 1e-3) / 1e9 / pk["hbm_gbs"],
            "peaks": pk["source"] + ", burst bf16 (kernel timed alone)",
            "timing": ("in-graph CUDA events around a train of %d launches over %d distinct queue copies (%.0f MB > L2), time / launches"
                       % (train["launches_per_train"], train["launches_per_train"], train["distinct_queue_bytes"] / 1e6))
                      if train else "CUDA events around a graph replay",
            "kernel_ms_single_launch": k_ms_single, "frac_single_launch": flops / (k_ms_single * 1e-3) / 1e12 / pk["bf16_tflops"],
            "l2": "flushed before every train; every launch of a train reads a queue copy that is not L2-resident"}
    prof = os.path.join(ROOT, "profiles", "r02_dram_traffic.json")
    if os.path.exists(prof):
        try:
            roof["traffic"] = json.load(open(prof)).get("infonce_tcx_kernel_dram_bytes")
        except (ValueError, OSError):
            pass
    secondary = None if args.no_secondary else secondary_measurements(dev, flush, pk)
    pretrain = None if args.no_secondary else pretrain_clips(0, 1, dev)

    affinity.restore(aff)
    cpu_rate, cpu_ms, cpu_n, cores = cpu_head_rate(60, 2, budget_s=15.0) if not args.no_cpu else (None, None, 0, 0)
    line = {
        "metric": METRIC, "value": 1e3 / ms_per_step, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": "moco_head_B256_K65536_d128", "B": B, "K": K, "d": D, "T": T, "queue_dtype": "bf16",
                   "algo": "tcgen05 single-pass (loss + dq in one queue sweep)",
                   "cuda_graph": steps_g[0].plan is None, "launch_plan": steps_g[0].plan is not None, "input_pool": POOL,
                   "queue_pool": QPOOL, "outputs": "loss, dq[256,128], top-1/top-5 hit counts (rank_gt = NULL), enqueue + pointer",
                   "l2": "no flush needed: consecutive steps use %d distinct queue replicas + workspaces (%.0f MB > 126 MB L2)"
                         % (QPOOL, QPOOL * (K * D * 2 + 74 * B * D * 4) / 1e6),
                   "timing": "EXACTLY `steps` steps (GraphedMoCoStep.step(): %s) back to back between ONE CUDA event pair on "
                             "the launch stream, synchronised on both sides; value = steps / elapsed"
                             % ("the step's three launches re-issued from a recorded launch plan (gca_plan_run), programmatic "
                                "dependent launch between them and across steps" if steps_g[0].plan is not None else
                                "one CUDA-graph launch per step")},
        "ms_per_step_isolated": ms_isolated, "isolated_note": "one step between its own event pair after a 256 MiB L2 flush "
                                "(round 1's method; includes the launch latency of a lone step)",
        "wall_s_total": wall, "loss_last": loss_last,
        "clocks": clocks,
        "cpu_affinity": {k: aff.get(k) for k in ("bound", "source", "cpus")},
        "e2e": {"value": 1e3 / e2e_ms, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms,
                "path": ("GraphedMoCoStep.step_host_io(): the step's first kernel reads q|k from the pinned host buffer over PCIe "
                         "(zero-copy, each once), its enqueue CTAs read all_k from it, " if zc_in else
                         "GraphedMoCoStep.step_host_io(): pinned host q|k|all_k -> H2D copy -> step (C ABI), ") +
                        "its last kernel stores loss|top-k hits|dq straight into the pinned host result buffer -- one CUDA-graph "
                        "launch (a synchronous step wants a single submission), " +
                        ("the device is synchronised and the loss read on the host every step" if e2e_sync else
                         "the host waits for the step's completion word in the pinned result buffer (gca_workspace_set_done_flag) "
                         "and reads the loss every step"),
                "zero_copy_in": zc_in,
                "loss_last": e2e_loss},
        "gpu_launches": launches,
        "roofline": roof,
        "secondary": secondary,
        "pretrain_clips_per_s": pretrain,
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
    from gca_b200 import affinity
    aff = affinity.bind_cpu_to_device(local_rank) if os.environ.get("GCA_BENCH_NO_BIND") != "1" else {"bound": False, "source": "off"}
    dist.init_process_group("nccl", device_id=dev)
    # QPOOL queue replicas per rank (consecutive timed steps stream disjoint working sets from HBM, see run_single); every
    # rank draws its own and rank 0's wins, as upstream (train_video_contrast_dis.py:233-242)
    torch.manual_seed(1 + 1000 * rank)
    mocos = [gca_b200.RGBMoCo(D, K=K, T=T, queue_dtype="bf16").to(dev) for _ in range(QPOOL)]
    for m in mocos:
        dist.broadcast(m.memory, 0)
    moco = mocos[0]
    pool = QPOOL
    batches = synthetic_batches(100 + rank, 4, B, 0, device=dev)
    states = [torch.tensor([0, 0], dtype=torch.int64, device=dev) for _ in range(QPOOL)]
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
        s = GraphedReplicaStep(mocos[i], B, state=states[i], exchange=exchange, fuse_exchange=(xmode == "fused"), want_rank=False)
        s.inputs[:2 * B].copy_(batches[i % len(batches)])
        s.prefer_graph = os.environ.get("GCA_BENCH_PREFER_GRAPH") == "1"     # A/B: one CUDA-graph launch per step
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
            s.moco.index = (s.moco.index + s.N) % K

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for i in range(args.warmup):
        one(i)
    if sampler:
        sampler.wait_first_sample()
    n0 = lib.gca_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    dist.barrier()
    if sampler:
        sampler.mark_start()
    t0 = time.perf_counter()
    ev0.record()
    for i in range(args.steps):                              # EXACTLY K steps, back to back, between one event pair per rank
        one(args.warmup + i)
    ev1.record()
    torch.cuda.synchronize()
    dist.barrier()
    wall = time.perf_counter() - t0
    if sampler:
        sampler.mark_end()
    clocks = sampler.stop() if sampler else None
    if graphed and steps_g[0].plan is not None and not steps_g[0].prefer_graph:
        launches = steps_g[0].plan.launches * args.steps     # the timed steps were re-issued from the launch plan
    else:
        launches = (steps_g[0].launches_per_step * args.steps) if graphed else int(lib.gca_launch_count() - n0)
    tot = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    dist.all_reduce(tot, op=dist.ReduceOp.MAX)              # max over ranks of the device-timed total
    ms_per_step = float(tot) / args.steps
    loss_last = float(steps_g[(args.warmup + args.steps - 1) % pool].loss)
    # round 1's figure for continuity: one step between its own event pair after an L2 flush, max over ranks (it contains
    # the launch skew between the ranks: a rank whose peers start the step later waits for their keys)
    iso = []
    for i in range(40):
        flush.fill_(i & 1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        one(i)
        b.record()
        torch.cuda.synchronize()
        if i >= 8:
            iso.append(a.elapsed_time(b))
    iso_t = torch.tensor([sum(iso) / len(iso)], device=dev)
    dist.all_reduce(iso_t, op=dist.ReduceOp.MAX)
    # replicas must stay bit-identical: compare a checksum of every queue replica and its ring pointer across ranks
    chk = torch.stack([torch.stack([m.memory.float().sum().double() for m in mocos]).sum(),
                       torch.stack([st_[0].double() for st_ in states]).sum()])
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
    sharded = sharded_strong = None
    if not args.no_sharded:
        sharded = time_sharded_k1m(rank, world, dev, flush)
        if B % world == 0:
            sharded_strong = time_sharded_k1m(rank, world, dev, flush, steps=100, rows_per_gpu=B // world)
    pretrain = None if args.no_sharded else pretrain_clips(rank, world, dev)
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
                       "cuda_graph": graphed and (steps_g[0].plan is None or steps_g[0].prefer_graph),
                       "launch_plan": graphed and steps_g[0].plan is not None and not steps_g[0].prefer_graph,
                       "queue_pool": QPOOL,
                       "outputs": "loss, dq[256,128], top-1/top-5 hit counts (rank_gt = NULL), enqueue + pointer",
                       "l2": "no flush needed: consecutive steps use %d distinct queue replicas + workspaces per rank (> 126 MB L2)" % QPOOL,
                       "timing": "EXACTLY `steps` steps back to back between ONE CUDA event pair per rank, barrier + synchronise on "
                                 "both sides, max over ranks; value = n_gpus * steps / elapsed (each global step processes "
                                 "n_gpus x 256 rows)"},
            "ms_per_step_isolated": float(iso_t), "wall_s_total": wall, "loss_last": loss_last, "replicas_consistent": consistent, "clocks": clocks,
            "cpu_affinity": {k: aff.get(k) for k in ("bound", "source", "cpus")},
            "e2e": {"value": world * 1e3 / float(e2e), "unit": UNIT, "h2d_bytes_per_step": 2 * B * D * 4,
                    "d2h_bytes_per_step": g0.outputs.numel() * 4, "ms_per_step": float(e2e),
                    "path": "GraphedReplicaStep.step_host_io(): one graph launch per step, q|k read from / results stored to pinned "
                            "host buffers by the step's own kernels, host barrier + stream sync every step" if host_io else
                            "H2D copy -> step -> D2H copy, host barrier + stream sync every step"},
            "gpu_launches": launches,
            "sharded_k1m": sharded,
            "sharded_k1m_strong": sharded_strong,
            "pretrain_clips_per_s": pretrain,
        }
        print(json.dumps(line), flush=True)
    # Leave without tearing the communicator down: destroy_process_group() after NCCL work was captured into CUDA graphs
    # has been seen to block forever (round 1, 2 GPUs); every rank is past its last collective here.
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def pretrain_clips(rank, world, dev):
    """BASELINE's second headline figure, pre-train clips/s (config 2: visual_moco.yaml shape, 64 videos/GPU x 2 clips of
    3x16x112x112, queue 65536, bf16 autocast) with the B200 head + one-launch EMA inside the trainer's step
    (tools/pretrain_step.py: ShuffleBN all-to-all + DDP at N > 1).  The encoder is the reference's R3D-18 re-stated layer for layer
    on cuDNN (same parameter shapes, checked on the CPU against the reference's class; random init -- the reference's own file is
    not on the GPU box and the backbone is outside the hot path): >97 % of this step is library code."""
    import types
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import pretrain_step as ps
    args = types.SimpleNamespace(batch=64, K=K, steps=8, warmup=4, quiet=True)
    torch.backends.cudnn.benchmark = True
    try:
        r = ps.run("b200", args, rank, world, dev)
        torch.cuda.empty_cache()
        return {"clips_per_s": r["clips_per_s"], "ms_per_step": r["ms_per_step"], "videos_per_gpu": 64, "n_gpus": world,
                "head_fwd_ms_median": r["head_fwd_ms_median"], "ema_ms": r["ema_ms"],
                "encoder": "the reference's R3D-18 re-stated (backbone_3d/resnet.py: identical layer shapes, 33.2 M parameters; MLP head) on cuDNN, bf16 autocast, "
                           "random init; synthetic Kinetics-shaped clips"}
    except Exception as e:                                   # a secondary figure must never take the headline down
        return {"error": "%s: %s" % (type(e).__name__, e)}


_KEEP = []


def time_sharded_k1m(rank, world, dev, flush, steps=200, warmup=10, rows_per_gpu=B):
    """BASELINE config 4: queue 2^20 x 128 (bf16) split along K over the ranks, `rows_per_gpu` query rows per GPU (256: the
    weak-scaling series, constant work per GPU; 256 / world: the strong-scaling series with B_global = 256);
    per-shard online-softmax partials merged with NCCL.  One CUDA graph per step (gca_b200.graphed.GraphedShardedStep);
    the first step is cross-checked against the eager ShardedRGBMoCo path (tests/sharded_graph_worker.py checks both against
    the oracle).  Per-step device time with an L2 flush before every step, max over ranks."""
    import torch
    import torch.distributed as dist
    import gca_b200
    from gca_b200.dist import ShardedRGBMoCo
    from gca_b200.graphed import GraphedShardedStep
    K1 = 1 << 20
    Bl = int(rows_per_gpu)
    torch.manual_seed(1)
    moco = ShardedRGBMoCo(D, K=K1, T=T, queue_dtype="bf16", device=dev)
    crit = gca_b200.NCESoftmaxLoss()
    batches = synthetic_batches(300 + rank, 2, Bl, 0, device=dev)
    # eager reference step on a copy of the shard
    shard0 = moco.memory.clone()
    q = batches[0][:Bl].clone().requires_grad_(True)
    out, _ = moco(q, batches[0][Bl:2 * Bl])
    loss_eager = crit(out)
    loss_eager.backward()
    torch.cuda.synchronize()
    mem_eager = moco.memory.clone()
    from gca_b200.peer import PeerShardLink
    link = PeerShardLink(Bl, D, device=dev)
    res = {}
    for variant in ("peer", "nccl"):
        moco.memory.copy_(shard0)
        moco.index = 0
        gs = GraphedShardedStep(moco, Bl, link=link if variant == "peer" else None).capture()
        gs.step(batches[0][:Bl], batches[0][Bl:2 * Bl])
        torch.cuda.synchronize()
        le = loss_eager.detach().reshape(1)
        if variant == "nccl":     # same kernels, same order: bit-identical to the eager path
            same = bool(torch.equal(gs.loss, le) and torch.equal(gs.dq, q.grad) and torch.equal(moco.memory, mem_eager))
        else:                     # the cross-rank merge has another (fixed) association order
            same = bool(abs(float(gs.loss) - float(le)) <= 2e-6 * abs(float(le))
                        and float((gs.dq - q.grad).abs().max()) <= 2e-5 * float(q.grad.abs().max())
                        and torch.equal(moco.memory, mem_eager))
        ev = []
        sync_word = torch.zeros(1, device=dev)
        for i in range(warmup + steps):
            flush.fill_(i & 1)
            pk = batches[i % 2]
            gs.q.copy_(pk[:Bl])
            gs.k.copy_(pk[Bl:2 * Bl])
            # line the ranks' streams up (a 4-byte all-reduce completes within a few us of its peers) so that the step time
            # does not include waiting for a peer that is still flushing its L2
            dist.all_reduce(sync_word)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            gs.step()
            b.record()
            if i >= warmup:
                ev.append((a, b))
        torch.cuda.synchronize()
        if variant == "peer":
            link.check()
        tot = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], device=dev)
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        ok = torch.tensor([1 if same else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        res[variant] = (float(tot) / steps, gs.launches_per_step, bool(int(ok)))
        _KEEP.append(gs)        # captured NCCL graphs are never torn down (see tests/sharded_graph_worker.py)
    ms, launches, ok_peer = res["peer"]
    flops = 4.0 * (Bl * world) * (K1 / world) * D
    return {"K": K1, "rows_per_gpu": Bl, "rows_global": Bl * world, "shard_rows": K1 // world, "ms_per_step": ms,
            "global_rows_per_s": Bl * world / ms * 1e3, "per_gpu_tflops": flops / (ms * 1e-3) / 1e12, "cuda_graph": True,
            "exchange": "NVLink peer memory inside the step's own launches (gca_shard_step_peer): q|k gather + cross-rank merge",
            "launches_per_step": launches, "collectives_per_step": 0, "matches_eager_path": ok_peer,
            "nccl_variant": {"ms_per_step": res["nccl"][0], "launches_per_step": res["nccl"][1], "collectives_per_step": 3,
                             "graph_equals_eager_path": res["nccl"][2]}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the extra K-sharded 2^20-row measurement")
    ap.add_argument("--no-secondary", action="store_true", help="N = 1: skip the secondary figures (fp32 mode, K = 2^20, graph head, retrieval)")
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
