#!/usr/bin/env python
"""Secondary kernels of the hot path, timed alone (CUDA events, L2 flushed between launches, mean of 30 after 5 warm-ups) with
their algorithmic bytes / FLOPs, next to the same op in eager PyTorch on the same GPU.  Prints JSON lines;
the summary goes to profiles/.  Not the headline benchmark (that is bench.py)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "video-graph-ssl_b200")):
    sys.path.insert(0, p)
import torch
import torch.nn.functional as F
from gca_b200 import functional as GF
import gca_b200
from oracle import graph as og

HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, n=30, warm=5, graph=False):
    """Mean device time of fn(); the work is captured in a CUDA graph when possible so that Python / launch overhead of
    multi-launch ops does not pollute the number (the same treatment for our ops and for eager torch)."""
    if graph:
        try:
            fn(); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            run = g.replay
        except Exception:
            torch.cuda.synchronize()
            run = fn
    else:
        run = fn
    return _time(run, n, warm)


def _time(fn, n, warm):
    ts = []
    for i in range(n + warm):
        flush.fill_(i & 1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(a.elapsed_time(b) * 1e3)
    return sum(ts) / len(ts)


def report(name, us, nbytes=None, flops=None, ref_us=None, **kw):
    r = {"kernel": name, "us": round(us, 2)}
    if nbytes:
        r.update(alg_MB=round(nbytes / 1e6, 2), GBps=round(nbytes / us / 1e3, 1), hbm_frac=round(nbytes / us / 1e3 / HBM, 3))
    if flops:
        r.update(alg_GFLOP=round(flops / 1e9, 2), TFLOPs=round(flops / us / 1e6, 1))
    if ref_us:
        r.update(torch_eager_us=round(ref_us, 2), speedup_vs_eager=round(ref_us / us, 1))
    r.update(kw)
    print(json.dumps(r), flush=True)


torch.manual_seed(0)
# ---- graph head: config 3 shapes
for shape, sub, tag in (((128, 192, 8, 14, 14), True, "c3-fmap"), ((128, 1024, 8, 1, 1), False, "c3-emb"), ((8, 128, 4, 1, 1), False, "c1")):
    Bv, C, T, H, W = shape
    x = torch.randn(*shape, device="cuda")
    wq = torch.randn(C // 2, C, 1, 1, 1, device="cuda") * (0.3 / (C * H * W) ** 0.5)
    wk = torch.randn(C // 2, C, 1, 1, 1, device="cuda") * (0.3 / C ** 0.5)
    wg = torch.randn(C, C, 1, 1, 1, device="cuda") / C ** 0.5
    gq, gk, sup = og._project(x, wq, sub, True), og._project(x, wk, sub, True), F.conv3d(x, wg)
    u = torch.rand(Bv, T, T, device="cuda")
    dy = torch.randn_like(sup)
    Dq = gq[0, :, 0].numel()
    fwd_bytes = 4 * Bv * (2 * T * Dq + 2 * C * T * H * W + 2 * T * T)
    bwd_bytes = 4 * Bv * (4 * T * Dq + 3 * C * T * H * W + 4 * T * T)          # read gq,gk,sup,dy + write d_gq,d_gk,d_sup
    from gca_b200 import _lib
    import ctypes
    P = _lib.ptr
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    gq, gk, sup = gq.contiguous(), gk.contiguous(), sup.contiguous()
    S, HW, Cq = gq[0, 0, 0].numel(), H * W, C // 2
    sim = torch.empty(Bv, T, T, device="cuda"); adj = torch.empty_like(sim); s_ = torch.empty_like(sim)
    y = torch.empty_like(sup); d_gq = torch.empty_like(gq); d_gk = torch.empty_like(gk); d_sup = torch.empty_like(sup)
    ws = torch.zeros(int(_lib.load().gca_graph_workspace_bytes(Bv, T)), dtype=torch.uint8, device="cuda")
    out = {}

    def fwd():          # straight through the C ABI (no autograd bookkeeping in the timed region)
        _lib.call("gca_graph_fwd", P(gq), P(gk), Cq, S, P(sup), C, HW, T, Bv, P(u), 0.5, 3, 1.0, 0, P(sim), P(adj), P(s_), P(y),
                  P(ws), ws.numel(), st)

    def bwd():
        _lib.call("gca_graph_bwd", P(gq), P(gk), Cq, S, P(sup), C, HW, T, Bv, P(sim), P(adj), P(s_), P(dy), 0.5, 3, 1.0, 0,
                  P(d_gq), P(d_gk), P(d_sup), P(ws), ws.numel(), st)

    wmat = og.edge_weight_matrix(T, 3, 0.5).cuda()

    def eager_fwd():        # the reference's op sequence (temporal_graph.py:161-239) in eager torch on the GPU
        Gq = gq.transpose(2, 1).contiguous().view(Bv, T, -1)
        Gk = gk.transpose(2, 1).contiguous().view(Bv, T, -1)
        sim_ = F.softmax(torch.matmul(Gq, Gk.permute(0, 2, 1)), dim=-1)
        adj_ = torch.zeros_like(sim_)
        hop = (torch.arange(T, device="cuda")[:, None] - torch.arange(T, device="cuda")[None, :]).abs()
        for h in range(4):
            adj_[:, hop == h] = sim_[:, hop == h] * float(wmat[0, h] if h < T else 0.0)
        eps = torch.finfo(torch.float32).eps          # rsample's op sequence with the uniforms given
        pr, uu = adj_.clamp(min=eps, max=1 - eps), u.clamp(min=eps, max=1 - eps)
        sg = torch.sigmoid((uu.log() - (-uu).log1p() + pr.log() - (-pr).log1p()) / 1.0)
        out["ye"] = torch.einsum('bij,bcjhw->bcihw', sg, sup) + sup

    t_f = timeit(fwd)
    t_e = timeit(eager_fwd)
    report("graph_fwd %s %s" % (tag, list(shape)), t_f, fwd_bytes, ref_us=t_e)
    t_b = timeit(bwd)
    report("graph_bwd %s %s" % (tag, list(shape)), t_b, bwd_bytes)

    def eager_fwd_bwd():    # the same op sequence through autograd (what the reference trainer runs per step)
        a, b, c = gq.detach().requires_grad_(True), gk.detach().requires_grad_(True), sup.detach().requires_grad_(True)
        Gq = a.transpose(2, 1).contiguous().view(Bv, T, -1)
        Gk = b.transpose(2, 1).contiguous().view(Bv, T, -1)
        sim_ = F.softmax(torch.matmul(Gq, Gk.permute(0, 2, 1)), dim=-1)
        hop = (torch.arange(T, device="cuda")[:, None] - torch.arange(T, device="cuda")[None, :]).abs()
        wfull = torch.zeros(T, T, device="cuda")
        for h in range(min(4, T)):
            wfull[hop == h] = float(wmat[0, h])
        adj_ = sim_ * wfull
        eps = torch.finfo(torch.float32).eps
        pr, uu = adj_.clamp(min=eps, max=1 - eps), u.clamp(min=eps, max=1 - eps)
        sg = torch.sigmoid((uu.log() - (-uu).log1p() + pr.log() - (-pr).log1p()) / 1.0)
        ye = torch.einsum('bij,bcjhw->bcihw', sg, c) + c
        ye.backward(dy)

    t_eb = timeit(eager_fwd_bwd, n=10, warm=3)
    report("graph core fwd+bwd %s: %d videos per step" % (tag, Bv), t_f + t_b, fwd_bytes + bwd_bytes, ref_us=t_eb,
           videos_per_s=round(Bv / ((t_f + t_b) * 1e-6), 0))

# ---- enqueue, negcos
moco = gca_b200.RGBMoCo(128, K=65536, queue_dtype="bf16").cuda()
keys = F.normalize(torch.randn(256, 128, device="cuda"))
report("enqueue N=256 bf16", timeit(lambda: GF.enqueue_(moco.memory, keys, 100)), 256 * 128 * 6)
p, z = torch.randn(128, 1024, device="cuda", requires_grad=True), torch.randn(128, 1024, device="cuda")
report("negcos fwd+grad [128,1024]", timeit(lambda: GF.neg_cosine(p, z)), 12 * 128 * 1024,
       ref_us=timeit(lambda: torch.autograd.grad(-F.cosine_similarity(p, z, dim=-1).mean(), p)))

# ---- InfoNCE fp32 parity mode (CUDA-core kernel) and eager torch on the GPU (what the reference would run)
K, B = 65536, 256
mem32 = F.normalize(torch.randn(K, 128, device="cuda"))
q, k = F.normalize(torch.randn(B, 128, device="cuda")), F.normalize(torch.randn(B, 128, device="cuda"))
report("infonce fp32 (ffma) fwd+grad B=256 K=65536", timeit(lambda: GF.infonce_forward(q, k, mem32, 0.07, algo="ffma")), flops=4.0 * B * K * 128)


def eager_head():
    qq = q.clone().requires_grad_(True)
    queue = mem32.clone().detach()
    lg = torch.cat(((qq * k).sum(1, keepdim=True), torch.mm(queue, qq.t()).t()), 1) / 0.07
    loss = F.cross_entropy(lg, torch.zeros(B, dtype=torch.long, device="cuda"))
    loss.backward()
    lg.detach().topk(5, 1, True, True)
    mem32.index_copy_(0, torch.arange(B, device="cuda"), k)


report("reference ops in eager torch on the GPU (fp32 head step)", timeit(eager_head), flops=4.0 * B * K * 128)
for Bq in (64, 256):
    qb, kb = q[:Bq].contiguous(), k[:Bq].contiguous()
    report("infonce bf16 (tcgen05) fused fwd+grad+finalize B=%d K=65536" % Bq,
           timeit(lambda: GF.infonce_forward(qb, kb, moco.memory, 0.07, algo="tcgen05")), K * 128 * 2, flops=4.0 * Bq * K * 128)
# a "learnt" encoder: every positive beats the whole queue, so the rank count of each step is skipped (warp-uniform branch)
report("infonce bf16 (tcgen05) fused fwd+grad+finalize B=256 K=65536, positives dominate (k = q)",
       timeit(lambda: GF.infonce_forward(q, q, moco.memory, 0.07, algo="tcgen05")), K * 128 * 2, flops=4.0 * B * K * 128)
# projection-head tail fused (gca_moco_step_proj) vs F.normalize x2 + step + enqueue + the normalisation's autograd backward
zq_, zk_ = torch.randn(B, 128, device="cuda") * 3, torch.randn(B, 128, device="cuda")
def proj_fused():
    GF.moco_step_proj(zq_, zk_, moco.memory, 0.07, 0)
def proj_unfused():
    z = zq_.detach().requires_grad_(True)
    qn, kn = F.normalize(z), F.normalize(zk_)
    o = GF.infonce_forward(qn, kn, moco.memory, 0.07, algo="tcgen05")
    GF.enqueue_(moco.memory, kn, 0)
    qn.backward(o["dq_unit"])
report("head step from un-normalised projections, normalisation fused (gca_moco_step_proj) B=256 K=65536", timeit(proj_fused),
       K * 128 * 2, flops=4.0 * B * K * 128, ref_us=timeit(proj_unfused))
K1 = 1 << 20
big = F.normalize(torch.randn(K1, 128, device="cuda")).to(torch.bfloat16)
for Bq in (64, 256):
    qb, kb = q[:Bq].contiguous(), k[:Bq].contiguous()
    report("infonce bf16 (tcgen05) fused fwd+grad+finalize B=%d K=2^20" % Bq,
           timeit(lambda: GF.infonce_forward(qb, kb, big, 0.07, algo="tcgen05"), n=10), K1 * 128 * 2, flops=4.0 * Bq * K1 * 128)

# ---- EMA of an R3D-18-sized encoder (33.4 M fp32 parameters in 62 tensors)
from gca_b200.ema import MomentumUpdater
shapes = [(64, 3, 7, 7, 7)] + [(c, c, 3, 3, 3) for c in (64,) * 4 + (128,) * 4 + (256,) * 4 + (512,) * 4] + [(512, 512, 3, 3, 3)] * 2
class Bag(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.ps = torch.nn.ParameterList([torch.nn.Parameter(torch.randn(*sh, device="cuda") * 0.01) for sh in shapes])
ma, mb = Bag(), Bag()
upd = MomentumUpdater(ma, mb)
def eager_ema():
    for p1, p2 in zip(ma.parameters(), mb.parameters()):
        p2.data.mul_(0.999).add_(p1.detach().data, alpha=0.001)
report("ema update %.1f M params (1 launch)" % (upd.numel / 1e6), timeit(lambda: upd.step(0.999)), 12 * upd.numel, ref_us=timeit(eager_ema))

# ---- retrieval (config 5)
gal, qry = torch.randn(13320, 512, device="cuda"), torch.randn(3783, 512, device="cuda")
report("sim_topk 3783x13320x512 k=50", timeit(lambda: GF.cosine_topk(qry, gal, 50), n=5, warm=2), flops=2.0 * 3783 * 13320 * 512,
       ref_us=timeit(lambda: (F.normalize(qry) @ F.normalize(gal).t()).topk(50, 1), n=5, warm=2))
