"""GPU parity tests proper: every C-ABI entry point against the CPU oracle / the golden fixtures generated from the
reference, on the same seeded inputs.  Integer and index work is compared bit-exactly; floating point within the
tolerances BASELINE.json states: loss 1e-5 relative (fp32 mode) / 2e-3 (bf16 mode), gradients 1e-2 relative.
Run with `pytest -m gpu` on a B200."""
import os
import sys
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle
from oracle import graph as og

pytestmark = pytest.mark.gpu

LOSS_RTOL_FP32 = 1e-5
LOSS_RTOL_BF16 = 2e-3
GRAD_RTOL = 1e-2
# fp32 mode at d == 128 (GCA_ALGO_TC32): fp32-grade logits on the tensor cores (six bf16 piece products), the gradient
# accumulates two pieces of each factor (error ~2^-17): the same bar as the CUDA-core kernel (algo="ffma")
GRAD_RTOL_FP32_TC = 1e-4


def T_(a):
    return torch.from_numpy(np.asarray(a))


def cu(t):
    return t.cuda()


def bf16r(t):
    return t.to(torch.bfloat16).float()


def tcq(q, T):
    """The queries as the tcgen05 kernels see them: log2(e)/T is folded into the bf16 query block, so it is q * log2(e)/T
    (an fp32 product) that is rounded to bf16 -- expressed back in q units (float64)."""
    c2 = torch.tensor(1.0 / T, dtype=torch.float32) * torch.tensor(1.4426950408889634, dtype=torch.float32)
    return (q.float() * c2).to(torch.bfloat16).double() / c2.double()


def rel_max(a, b):
    """max |a - b| / max |b|  (norm-wise relative error)."""
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture(scope="module")
def GF(lib):
    import gca_b200
    from gca_b200 import functional
    return functional


def unit_rows(n, d, gen):
    return F.normalize(torch.randn(n, d, generator=gen))


# ============================================================================================ K3 enqueue
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("K,d,N,index", [(256, 128, 8, 0), (256, 128, 40, 236), (256, 128, 256, 17), (4096, 128, 64, 4086),
                                         (96, 32, 96, 95), (65536, 128, 512, 65400), (64, 1024, 3, 62)])
def test_enqueue_bit_exact(GF, dtype, K, d, N, index):
    gen = torch.Generator().manual_seed(K + N + index)
    mem = torch.randn(K, d, generator=gen).to(dtype)
    keys = torch.randn(N, d, generator=gen)
    ref = mem.clone()
    new_ref = oracle.enqueue(ref, keys, index)
    dev = cu(mem)
    new_idx = GF.enqueue_(dev, cu(keys), index)
    assert new_idx == new_ref                                        # pointer: exact
    assert torch.equal(dev.cpu(), ref)                               # slot -> row mapping and contents: exact


def test_enqueue_golden_wrap(GF, golden):
    g = golden("infonce_wrap")
    dev = cu(T_(g["memory_before"]).clone())
    idx = int(g["start_index"])
    for st in range(int(g["steps"])):
        idx = GF.enqueue_(dev, cu(T_(g[f"all_k{st}"])), idx)
        assert idx == int(g[f"index_after{st}"])
    assert torch.equal(dev.cpu(), T_(g["memory_after"]))


@pytest.mark.parametrize("W", [2, 4, 8])
def test_enqueue_sharded_ownership(GF, W):
    gen = torch.Generator().manual_seed(W)
    K, d, N, index = 1024, 128, 200, 1024 - 70
    full = torch.randn(K, d, generator=gen).to(torch.bfloat16)
    keys = torch.randn(N, d, generator=gen)
    ref = full.clone()
    oracle.enqueue(ref, keys, index)
    Ks = K // W
    shards = [cu(full[r * Ks:(r + 1) * Ks].clone()) for r in range(W)]
    for r in range(W):
        assert GF.enqueue_(shards[r], cu(keys), index, K_global=K, k_begin=r * Ks) == (index + N) % K
    assert torch.equal(torch.cat([s.cpu() for s in shards]), ref)


# ============================================================================================ K1/K2 InfoNCE, fp32 parity mode
def check_infonce(GF, q, k, mem, T, algo, loss_rtol, grad_rtol, ref_q=None, ref_mem=None, check_logits=False):
    """Run the fused head on the GPU and compare with the fp64 oracle on (ref_q, ref_mem) (default: same inputs)."""
    rq = (q if ref_q is None else ref_q).double()
    rm = (mem.float() if ref_mem is None else ref_mem).double()
    o = oracle.infonce_step(rq, k.double(), rm.clone(), 0, T)
    r = GF.infonce_forward(cu(q), cu(k), cu(mem), T, algo=algo, want_grad=True, materialize=check_logits)
    torch.cuda.synchronize()
    loss = float(r["loss"])
    assert abs(loss - float(o["loss"])) <= loss_rtol * abs(float(o["loss"])), (loss, float(o["loss"]))
    assert rel_max(r["lse"], o["lse"]) <= loss_rtol
    assert rel_max(r["loss_rows"], o["loss_rows"]) <= loss_rtol * 10
    assert rel_max(r["dq_unit"], o["dq"]) <= grad_rtol, rel_max(r["dq_unit"], o["dq"])
    if check_logits:
        lg = oracle.logits_full(rq, k.double(), rm, T)
        assert rel_max(r["logits"], lg) <= 1e-5
    return r, o


def test_infonce_fp32_golden_small(GF, golden):
    g = golden("infonce_small")
    T = float(g["T"])
    mem = T_(g["memory_before"])
    q, k = T_(g["q0"]), T_(g["k0"])
    r, _ = check_infonce(GF, q, k, mem, T, "ffma", LOSS_RTOL_FP32, 1e-4, check_logits=True)
    assert abs(float(r["loss"]) - float(g["loss0"])) <= LOSS_RTOL_FP32 * float(g["loss0"])     # vs the reference itself
    np.testing.assert_allclose(r["dq_unit"].cpu().numpy(), g["dq0"], rtol=1e-3, atol=1e-7)
    np.testing.assert_array_equal(r["rank"].cpu().numpy(), g["rank0"])                           # integer: exact
    np.testing.assert_allclose(r["logits"][:, :9].cpu().numpy(), g["logits_head0"], rtol=1e-5, atol=1e-5)


def test_infonce_fp32_config1(GF, golden):
    """BASELINE config 1 / SURVEY Appendix C G1 (B=32, K=4096, d=128), inputs regenerated from the seed."""
    g = golden("infonce_c1")
    torch.manual_seed(1)
    mem = F.normalize(torch.randn(4096, 128))
    q = F.normalize(torch.randn(32, 128))
    k = F.normalize(torch.randn(32, 128))
    r, _ = check_infonce(GF, q, k, mem, 0.07, "ffma", LOSS_RTOL_FP32, 1e-4)
    assert abs(float(r["loss"]) - 8.9238758087) <= LOSS_RTOL_FP32 * 8.9238758087
    assert float(r["dq_unit"].abs().sum()) == pytest.approx(130.3457336426, rel=1e-4)
    np.testing.assert_array_equal(r["rank"].cpu().numpy(), g["rank0"])
    top1 = float((r["rank"] < 1).float().mean() * 100)
    top5 = float((r["rank"] < 5).float().mean() * 100)
    assert [top1, top5] == list(g["acc0"])


@pytest.mark.parametrize("B,K,d", [(1, 256, 128), (32, 4096, 128), (64, 1000, 128), (256, 4096, 128), (5, 192, 64),
                                   (130, 777, 32), (16, 512, 512), (9, 320, 1024), (256, 16384, 128)])
@pytest.mark.parametrize("qdtype", [torch.float32, torch.bfloat16])
def test_infonce_ffma_sweep(GF, B, K, d, qdtype):
    gen = torch.Generator().manual_seed(B * 7 + K + d)
    mem = unit_rows(K, d, gen).to(qdtype)
    q, k = unit_rows(B, d, gen), unit_rows(B, d, gen)
    r, o = check_infonce(GF, q, k, mem, 0.07, "ffma", LOSS_RTOL_FP32, 1e-4)
    # ranks are integers: exact wherever the margin to the nearest negative is above fp32 noise
    lg = oracle.logits_full(q.double(), k.double(), mem.double(), 0.07)
    margin = (lg[:, 1:] - lg[:, :1]).abs().min(dim=1).values
    ok = margin > 3e-5                                   # fp32 noise of a logit of magnitude <= 1/T
    assert ok.float().mean() >= 0.4
    assert torch.equal(r["rank"].cpu().long()[ok], o["rank"][ok])


@pytest.mark.parametrize("B,K", [(64, 1024), (130, 1000), (32, 4096), (256, 65536)])
def test_infonce_tc32_fp32_grade_on_tensor_cores(GF, B, K):
    """GCA_ALGO_TC32: the fp32 parity mode on tcgen05 (exact 3-way bf16 split of q and the fp32 queue, six piece products
    per logit).  Loss / lse within the fp32-mode bar of BASELINE.json (1e-5) against the fp64 oracle on the SAME fp32
    inputs, gradient within 1e-4 (two bf16 pieces of P and of the queue), rank exact away from fp32-noise ties; ragged K (tile tail crosses into
    the next plane of the split) and B (padding rows)."""
    torch.manual_seed(B + K)
    T = 0.07
    mem = F.normalize(torch.randn(K, 128))
    q, k = F.normalize(torch.randn(B, 128)), F.normalize(torch.randn(B, 128))
    r, o = check_infonce(GF, q, k, mem, T, "tc32", LOSS_RTOL_FP32, GRAD_RTOL)
    assert rel_max(r["dq_unit"], o["dq"]) <= 1e-4
    neg = oracle.logits_full(q.double(), k.double(), mem.double(), T)[:, 1:]
    pos = ((q.double() * k.double()).sum(1) / T)[:, None]
    ok = (neg - pos).abs().min(1).values > 3e-5
    assert int(ok.sum()) >= min(B // 2, 32)
    assert torch.equal(r["rank"].cpu().long()[ok], (neg > pos).sum(1)[ok])
    # same call through the ffma family: the two fp32 modes agree far inside the bar
    r2 = GF.infonce_forward(cu(q), cu(k), cu(mem), T, algo="ffma", want_grad=True)
    assert abs(float(r["loss"]) - float(r2["loss"])) <= 2e-6 * abs(float(r2["loss"]))


def test_infonce_tc32_unnormalised_inputs_take_the_exact_max_pass(GF):
    """Rows far from unit norm leave the fixed-reference window: the CTA redoes its keys with the exact row max."""
    torch.manual_seed(9)
    B, K, T = 96, 2048, 0.07
    mem = torch.randn(K, 128) * 1.5
    q, k = torch.randn(B, 128) * 2.0, torch.randn(B, 128)
    check_infonce(GF, q, k, mem, T, "tc32", LOSS_RTOL_FP32, GRAD_RTOL)


def test_infonce_ffma_unnormalised_inputs(GF):
    """Online max must hold for logits far outside [-1/T, 1/T]."""
    gen = torch.Generator().manual_seed(5)
    mem = torch.randn(2048, 128, generator=gen) * 3
    q, k = torch.randn(48, 128, generator=gen), torch.randn(48, 128, generator=gen)
    check_infonce(GF, q, k, mem, 0.5, "ffma", LOSS_RTOL_FP32, 1e-4)


# ============================================================================================ K1 tcgen05 (bf16 queue)
TC_SHAPES = [(128, 128), (1, 128), (64, 256), (128, 384), (200, 1000), (256, 4096), (256, 65536), (64, 65536), (32, 4096),
             (512, 8192)]


@pytest.mark.parametrize("B,K", TC_SHAPES)
def test_infonce_tcgen05_vs_bf16_input_oracle(GF, B, K):
    """The kernel's arithmetic: bf16-rounded q and queue, fp32 accumulate.  Against the fp64 oracle fed the SAME
    rounded inputs the loss must agree to ~1e-5 (only P is re-rounded, which touches the gradient alone)."""
    gen = torch.Generator().manual_seed(B + K)
    mem = unit_rows(K, 128, gen).to(torch.bfloat16)
    q, k = unit_rows(B, 128, gen), unit_rows(B, 128, gen)
    T = 0.07
    rq = tcq(q, T)
    o = oracle.infonce_step(rq.double(), k.double(), mem.double().clone(), 0, T)
    # the positive logit is taken from the unrounded q (fp32), the negatives from the rounded one
    pos = (q.double() * k.double()).sum(1) / T
    neg = (rq.double() @ mem.double().t()) / T
    lse = torch.logsumexp(torch.cat([pos[:, None], neg], 1), 1)
    loss_ref = float((lse - pos).mean())
    r = GF.infonce_forward(cu(q), cu(k), cu(mem), T, algo="tcgen05", want_grad=True)
    torch.cuda.synchronize()
    assert abs(float(r["loss"]) - loss_ref) <= 2e-5 * abs(loss_ref), (float(r["loss"]), loss_ref)
    assert rel_max(r["lse"], lse) <= 2e-5
    assert rel_max(r["pos"], pos) <= 1e-5
    p0 = torch.exp(pos - lse)
    dq_ref = ((p0 - 1)[:, None] * k.double() + torch.exp(neg - lse[:, None]) @ mem.double()) / (T * B)
    assert rel_max(r["dq_unit"], dq_ref) <= 5e-3, rel_max(r["dq_unit"], dq_ref)
    rank_ref = (neg > pos[:, None]).sum(1)
    margin = (neg - pos[:, None]).abs().min(dim=1).values
    ok = margin > 1e-3
    assert torch.equal(r["rank"].cpu().long()[ok], rank_ref[ok])
    del o


def test_infonce_tcgen05_rank_count_skipped_when_positive_dominates(GF):
    """Rows whose positive beats every negative let a warp skip the rank count of a step (warp-uniform branch): ranks stay
    exact for all-dominant rows, for mixed warps (first 40 rows dominant, the rest random) and hit counts follow."""
    gen = torch.Generator().manual_seed(21)
    T, B, K = 0.07, 192, 4096
    mem = unit_rows(K, 128, gen).to(torch.bfloat16)
    q = unit_rows(B, 128, gen)
    k = unit_rows(B, 128, gen)
    k[:40] = q[:40]                                                        # cos = 1: nothing in the queue can beat it
    k[128:160] = q[128:160]                                                # one whole warp (32 rows) dominant
    rq = tcq(q, T)
    pos = (q.double() * k.double()).sum(1) / T
    neg = (rq.double() @ mem.double().t()) / T
    rank_ref = (neg > pos[:, None]).sum(1)
    lse = torch.logsumexp(torch.cat([pos[:, None], neg], 1), 1)
    r = GF.infonce_forward(cu(q), cu(k), cu(mem), T, algo="tcgen05", want_grad=True)
    torch.cuda.synchronize()
    rank = r["rank"].cpu().long()
    assert int(rank[:40].sum()) == 0 and int(rank[128:160].sum()) == 0
    margin = (neg - pos[:, None]).abs().min(dim=1).values
    ok = margin > 1e-3
    assert torch.equal(rank[ok], rank_ref[ok])
    assert r["hits"].cpu().tolist() == [int((rank < 1).sum()), int((rank < 5).sum())]
    assert abs(float(r["loss"]) - float((lse - pos).mean())) <= 3e-5 * float((lse - pos).mean())


def test_infonce_tcgen05_saturated_positive_keeps_the_packed_loss_word_sane(GF):
    """A perfectly learnt batch (k = q, tiny temperature): every row loss is ~0 and may round to -1 ulp; the packed
    fixed-point loss word must not wrap -- loss stays ~0 and non-negative, hits = B."""
    gen = torch.Generator().manual_seed(33)
    B, K, T = 128, 2048, 0.02
    mem = unit_rows(K, 128, gen).to(torch.bfloat16)
    q = unit_rows(B, 128, gen)
    r = GF.infonce_forward(cu(q), cu(q.clone()), cu(mem), T, algo="tcgen05", want_grad=True)
    torch.cuda.synchronize()
    assert 0.0 <= float(r["loss"]) <= 1e-6
    assert r["hits"].cpu().tolist() == [B, B] and int(r["rank"].sum()) == 0


def test_infonce_tcgen05_unnormalised_inputs_leave_the_packed_loss_word(GF):
    """Logits far outside [-1/T, 1/T] do not fit the packed fixed-point loss word of the finalize kernel: the stream /
    prep kernels flag it and the launch takes the wide accumulators -- then the flag is cleared, so a unit-row step on the
    same workspace goes back to the packed word.  Loss mean and hit counts are checked in all three launches."""
    gen = torch.Generator().manual_seed(9)
    T = 0.07
    mem_big = (torch.randn(4096, 128, generator=gen) * 0.5).to(torch.bfloat16)          # |row| ~ 5.7: logits up to ~ +-400
    mem_unit = unit_rows(4096, 128, gen).to(torch.bfloat16)
    q_big, k_big = torch.randn(64, 128, generator=gen) * 0.5, torch.randn(64, 128, generator=gen) * 0.5
    q_u, k_u = unit_rows(64, 128, gen), unit_rows(64, 128, gen)
    for q, k, mem in ((q_u, k_u, mem_unit), (q_big, k_big, mem_big), (q_u, -k_u * 1.5, mem_unit), (q_u, k_u, mem_unit)):
        rq = tcq(q, T)
        pos = (q.double() * k.double()).sum(1) / T
        neg = (rq.double() @ mem.double().t()) / T
        lse = torch.logsumexp(torch.cat([pos[:, None], neg], 1), 1)
        loss_ref = float((lse - pos).mean())
        r = GF.infonce_forward(cu(q), cu(k), cu(mem), T, algo="tcgen05", want_grad=True)
        torch.cuda.synchronize()
        assert abs(float(r["loss"]) - loss_ref) <= 3e-5 * abs(loss_ref), (float(r["loss"]), loss_ref)
        assert abs(float(r["loss"]) - float(r["loss_rows"].double().mean())) <= 1e-6 * abs(loss_ref)
        rank = r["rank"].cpu()
        assert r["hits"].cpu().tolist() == [int((rank < 1).sum()), int((rank < 5).sum())]


def _rank_band_check(rank, neg, pos, band):
    """Ranks are integers, but a rank computed from fp32 tensor-core sums can only be pinned up to the negatives that sit
    within the accumulation noise of the positive: |rank - rank_ref| <= #{j : |neg_bj - pos_b| <= band} for every row.
    Returns (rows whose band is empty, mean band population)."""
    rank_ref = (neg > pos[:, None]).sum(1)
    inband = ((neg - pos[:, None]).abs() <= band).sum(1)
    diff = (rank.cpu().long() - rank_ref).abs()
    assert bool((diff <= inband).all()), (int((diff - inband).max()), int(inband.max()))
    exact_rows = inband == 0
    assert torch.equal(rank.cpu().long()[exact_rows], rank_ref[exact_rows])
    return int(exact_rows.sum()), float(inband.float().mean())


@pytest.mark.parametrize("algo,qdt", [("tcgen05", torch.bfloat16), ("ffma", torch.float32)])
def test_infonce_rank_and_hits_at_the_headline_shape(GF, algo, qdt):
    """(B, K) = (256, 65536): the rank of the positive in EVERY row against the fp64 oracle fed the kernel's own inputs,
    within the noise band of fp32 accumulation (4e-7 in dot-product units -- the 65536 negatives of a row are ~3e-6 apart
    around a random positive, so about one row in four has a negative inside the band),
    and the top-1 / top-5 hit counts the step reports.  Half of the positives are made hard (rank < 50) so that top-5 is
    not trivially empty."""
    gen = torch.Generator().manual_seed(4242)
    B, K, T = 256, 65536, 0.07
    mem = unit_rows(K, 128, gen).to(qdt)
    q = unit_rows(B, 128, gen)
    k = unit_rows(B, 128, gen)
    k[::2] = F.normalize(q[::2] * 0.42 + k[::2])                          # cos ~ 0.39: around the top few of 65536 negatives
    rq = tcq(q, T) if algo == "tcgen05" else q.double()
    pos = (q.double() * k.double()).sum(1) / T
    neg = (rq @ mem.double().t()) / T
    r = GF.infonce_forward(cu(q), cu(k), cu(mem), T, algo=algo, want_grad=True)
    torch.cuda.synchronize()
    exact, mean_band = _rank_band_check(r["rank"], neg, pos, 4e-7 / T)
    assert exact >= B // 2 and mean_band < 1.0, (exact, mean_band)       # the check is not vacuous
    rank_ref = (neg > pos[:, None]).sum(1)
    assert 0 < int((rank_ref < 5).sum()) < B
    rank = r["rank"].cpu().long()
    assert r["hits"].cpu().tolist() == [int((rank < 1).sum()), int((rank < 5).sum())]
    sure = ((neg - pos[:, None]).abs() <= 4e-7 / T).sum(1) == 0
    hits_ref = [int((rank_ref < 1).sum()), int((rank_ref < 5).sum())]
    slack = int((~sure).sum())
    assert all(abs(a - b) <= slack for a, b in zip(r["hits"].cpu().tolist(), hits_ref))


def test_infonce_tcgen05_top_hits_only_mode(lib, GF):
    """rank_gt == NULL (include/gca_b200.h): the caller wants the top-1 / top-5 hit counts only, so a warp may stop counting
    once all of its rows are past GCA_TOPK_RANK_CAP.  Hits, loss and gradient must equal the exact-rank call."""
    from gca_b200 import _lib
    from gca_b200._lib import ptr
    import ctypes
    gen = torch.Generator().manual_seed(77)
    B, K, T = 256, 16384, 0.07
    mem = cu(unit_rows(K, 128, gen).to(torch.bfloat16))
    q, k = unit_rows(B, 128, gen), unit_rows(B, 128, gen)
    k[::3] = F.normalize(q[::3] * 0.5 + k[::3])
    q, k = cu(q), cu(k)
    full = GF.infonce_forward(q, k, mem, T, algo="tcgen05", want_grad=True)
    ws = GF.workspace(q.device, GF.infonce_workspace_bytes(B, K, 128, 1, "tcgen05"), "hits_only")
    f32 = lambda *s: torch.empty(*s, dtype=torch.float32, device=q.device)
    loss, rows, lse, pos, dq = f32(1), f32(B), f32(B), f32(B), f32(B, 128)
    hits = torch.zeros(2, dtype=torch.int32, device=q.device)
    _lib.call("gca_infonce_fwd", ptr(q), ptr(k), ptr(mem), 1, B, K, 128, 1.0 / T, 2, ptr(loss), ptr(rows), ptr(lse), ptr(pos),
              None, ptr(hits), ptr(dq), None, ptr(ws), ws.numel(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert hits.cpu().tolist() == full["hits"].cpu().tolist() and 0 < int(hits[1]) < B
    assert torch.equal(loss.view(()), full["loss"].view(())) and torch.equal(dq, full["dq_unit"]) and torch.equal(lse, full["lse"])


@pytest.mark.parametrize("T", [0.011, 0.005])
def test_infonce_tcgen05_tiny_temperature(GF, T):
    """1/T beyond 60 log2 units: the first sweep's reference exponent is no longer 0 (T = 0.005: 1/T - 100 log2 units) and
    rows whose best key is far below 1/T leave the validity window, so their CTA redoes the range with the exact row max."""
    gen = torch.Generator().manual_seed(5)
    B, K = 128, 4096
    mem = unit_rows(K, 128, gen).to(torch.bfloat16)
    q, k = unit_rows(B, 128, gen), unit_rows(B, 128, gen)
    k[:32] = q[:32]
    rq = tcq(q, T)
    pos = (q.double() * k.double()).sum(1) / T
    neg = (rq @ mem.double().t()) / T
    lse = torch.logsumexp(torch.cat([pos[:, None], neg], 1), 1)
    dq_ref = ((torch.exp(pos - lse) - 1)[:, None] * k.double() + torch.exp(neg - lse[:, None]) @ mem.double()) / (T * B)
    r = GF.infonce_forward(cu(q), cu(k), cu(mem), T, algo="tcgen05", want_grad=True)
    torch.cuda.synchronize()
    assert rel_max(r["lse"], lse) <= 2e-5
    assert abs(float(r["loss"]) - float((lse - pos).mean())) <= 3e-5 * float((lse - pos).mean())
    assert rel_max(r["dq_unit"], dq_ref) <= 5e-3
    _rank_band_check(r["rank"], neg, pos, 1e-6 / T)


@pytest.mark.parametrize("B,K", [(32, 4096), (256, 65536)])
def test_infonce_bf16_mode_within_baseline_tolerance(GF, B, K):
    """bf16 mode against the fp32 reference arithmetic on UNrounded inputs: loss 2e-3, gradient 1e-2 (BASELINE.json)."""
    gen = torch.Generator().manual_seed(K - B)
    mem32 = unit_rows(K, 128, gen)
    q, k = unit_rows(B, 128, gen), unit_rows(B, 128, gen)
    check_infonce(GF, q, k, mem32.to(torch.bfloat16), 0.07, "tcgen05", LOSS_RTOL_BF16, GRAD_RTOL, ref_mem=mem32)


def test_infonce_tcgen05_materialised_logits(GF):
    gen = torch.Generator().manual_seed(77)
    B, K = 70, 700
    mem = unit_rows(K, 128, gen).to(torch.bfloat16)
    q, k = unit_rows(B, 128, gen), unit_rows(B, 128, gen)
    r = GF.infonce_forward(cu(q), cu(k), cu(mem), 0.07, algo="tcgen05", want_grad=False, materialize=True)
    neg = (tcq(q, 0.07) @ mem.double().t()) / 0.07
    assert rel_max(r["logits"][:, 1:], neg) <= 1e-5
    assert rel_max(r["logits"][:, 0], (q.double() * k.double()).sum(1) / 0.07) <= 1e-5


def test_infonce_tcgen05_unnormalised_inputs_trigger_rescale(GF):
    """Rows whose max grows by more than 2^8 between tiles exercise the lazy O-rescale path."""
    gen = torch.Generator().manual_seed(9)
    B, K = 128, 2048
    mem = torch.randn(K, 128, generator=gen)
    mem[1500:] *= 6.0                                   # later tiles carry much larger logits
    mem = mem.to(torch.bfloat16)
    q, k = torch.randn(B, 128, generator=gen), torch.randn(B, 128, generator=gen)
    T = 1.0
    rq = tcq(q, T)
    pos = (q.double() * k.double()).sum(1) / T
    neg = (rq.double() @ mem.double().t()) / T
    lse = torch.logsumexp(torch.cat([pos[:, None], neg], 1), 1)
    dq_ref = ((torch.exp(pos - lse) - 1)[:, None] * k.double() + torch.exp(neg - lse[:, None]) @ mem.double()) / (T * B)
    r = GF.infonce_forward(cu(q), cu(k), cu(mem), T, algo="tcgen05", want_grad=True)
    assert rel_max(r["lse"], lse) <= 2e-5
    assert rel_max(r["dq_unit"], dq_ref) <= 5e-3


@pytest.mark.parametrize("algo,qdtype", [("ffma", torch.float32), ("tcgen05", torch.bfloat16)])
def test_infonce_two_pass_backward_equals_single_pass(GF, algo, qdtype):
    gen = torch.Generator().manual_seed(31)
    B, K = 96, 3000
    mem = cu(unit_rows(K, 128, gen).to(qdtype))
    q, k = cu(unit_rows(B, 128, gen)), cu(unit_rows(B, 128, gen))
    r = GF.infonce_forward(q, k, mem, 0.07, algo=algo, want_grad=True)
    dq = GF.infonce_backward_recompute(q, k, mem, 0.07, r["lse"], 1.0 / B, algo=algo)
    assert rel_max(dq, r["dq_unit"]) <= (1e-5 if algo == "ffma" else 5e-3)


@pytest.mark.parametrize("algo,qdtype", [("ffma", torch.float32), ("tcgen05", torch.bfloat16)])
def test_infonce_deterministic(GF, algo, qdtype):
    gen = torch.Generator().manual_seed(2)
    mem = cu(unit_rows(8192, 128, gen).to(qdtype))
    q, k = cu(unit_rows(256, 128, gen)), cu(unit_rows(256, 128, gen))
    a = GF.infonce_forward(q, k, mem, 0.07, algo=algo)
    b = GF.infonce_forward(q, k, mem, 0.07, algo=algo)
    for key in ("loss", "lse", "dq_unit", "rank"):
        assert torch.equal(a[key], b[key]), key


# ============================================================================================ sharded path on one GPU
@pytest.mark.parametrize("W,algo,qdtype", [(2, "ffma", torch.float32), (8, "ffma", torch.float32),
                                           (4, "tcgen05", torch.bfloat16), (8, "tcgen05", torch.bfloat16)])
def test_shard_abi_merge_equals_unsharded(GF, W, algo, qdtype):
    """Size-independent property: K-shard partials + combine + finish == the unsharded fused head."""
    from gca_b200.dist import ShardCompute
    gen = torch.Generator().manual_seed(W)
    B, K = 64, 8192
    mem = cu(unit_rows(K, 128, gen).to(qdtype))
    q, k = cu(unit_rows(B, 128, gen)), cu(unit_rows(B, 128, gen))
    full = GF.infonce_forward(q, k, mem, 0.07, algo=algo)
    comp = ShardCompute()
    Ks = K // W
    parts = [comp.shard_fwd(q, k, mem[r * Ks:(r + 1) * Ks].contiguous(), 0.07, algo, True) for r in range(W)]
    all_stats = torch.stack([p[1] for p in parts])                    # [W, 3, B]  (what the all-gather delivers)
    acc_sum = torch.zeros(B, 128, device="cuda")
    for r in range(W):
        lse, loss_rows, rank_gt = comp.shard_combine(all_stats, r, parts[r][0], parts[r][2])
        acc_sum += parts[r][2]                                        # what the reduce-scatter delivers
    dq_unit, loss = comp.shard_finish(acc_sum, k, parts[0][0], lse, loss_rows, 0.07)
    tol = 1e-5 if algo == "ffma" else 1e-4
    assert rel_max(lse, full["lse"]) <= 1e-6
    assert abs(float(loss) - float(full["loss"])) <= 1e-6 * abs(float(full["loss"]))
    assert torch.equal(rank_gt, full["rank"])
    assert rel_max(dq_unit, full["dq_unit"]) <= tol


# ============================================================================================ drop-in modules
@pytest.mark.parametrize("queue_dtype", ["fp32", "bf16"])
def test_rgbmoco_module_steps_like_reference(lib, golden, queue_dtype):
    """Three consecutive trainer-style steps (train_video_contrast_dis.py:411-428) against the reference outputs."""
    import gca_b200
    g = golden("infonce_small")
    moco = gca_b200.RGBMoCo(128, K=256, T=float(g["T"]), queue_dtype=queue_dtype).cuda()
    moco.load_state_dict({"memory": T_(g["memory_before"])})
    crit = gca_b200.NCESoftmaxLoss()
    ltol, gtol = (LOSS_RTOL_FP32, GRAD_RTOL_FP32_TC) if queue_dtype == "fp32" else (LOSS_RTOL_BF16, GRAD_RTOL)
    for st in range(int(g["steps"])):
        q = cu(T_(g[f"q{st}"])).requires_grad_(True)
        k = cu(T_(g[f"k{st}"]))
        out, labels = moco(q, k)
        loss = crit(out)
        loss.backward()
        assert labels.dtype == torch.long and labels.shape == (8,) and int(labels.abs().sum()) == 0
        assert out.shape[0] == 8 and out.shape[1] == 257
        assert abs(float(loss) - float(g[f"loss{st}"])) <= ltol * float(g[f"loss{st}"])
        assert rel_max(q.grad, T_(g[f"dq{st}"])) <= gtol
        assert moco.index == int(g[f"index_after{st}"])                # pointer: exact
        _, pred = out.detach().topk(5, 1, True, True)                  # what `accuracy` does
        correct = pred.t().eq(labels.view(1, -1).expand_as(pred.t()))
        acc = [float(correct[:kk].reshape(-1).float().sum() * (100.0 / 8)) for kk in (1, 5)]
        if queue_dtype == "fp32":
            assert acc == list(g[f"acc{st}"])
    mem_after = moco.state_dict()["memory"].cpu()
    if queue_dtype == "fp32":
        assert torch.equal(mem_after, T_(g["memory_after"]))           # slot contents: exact
    else:
        assert torch.equal(mem_after, bf16r(T_(g["memory_after"])))    # exact RNE rounding of the same rows


def test_rgbmoco_all_k_and_jig(lib, golden):
    import gca_b200
    g = golden("infonce_wrap")
    moco = gca_b200.RGBMoCo(128, K=256, T=float(g["T"])).cuda()
    moco.load_state_dict({"memory": T_(g["memory_before"])})
    moco.index = int(g["start_index"])
    q = cu(T_(g["q0"])).requires_grad_(True)
    out, out_jig, labels = moco(q, cu(T_(g["k0"])), q_jig=q.detach().flip(0), all_k=cu(T_(g["all_k0"])))
    assert abs(float(out.loss) - float(g["loss0"])) <= LOSS_RTOL_FP32 * float(g["loss0"])
    assert moco.index == int(g["index_after0"]) and out_jig.shape == out.shape


@pytest.mark.parametrize("queue_dtype", ["fp32", "bf16"])
def test_rgbmoco_jig_head_against_reference_fixture(lib, golden, queue_dtype):
    """`RGBMoCo.forward(q, k, q_jig=...)` (mem_moco.py:73-76, 85-86): BOTH heads' losses and gradients against the values
    the reference's own module + NCESoftmaxLoss produced (oracle/gen_golden_cmc.py)."""
    import gca_b200
    j = golden("rgb_jig")
    moco = gca_b200.RGBMoCo(128, K=int(j["K"]), T=float(j["T"]), queue_dtype=queue_dtype).cuda()
    moco.load_state_dict({"memory": T_(j["memory_before"])})
    crit = gca_b200.NCESoftmaxLoss()
    ltol, gtol = (LOSS_RTOL_FP32, GRAD_RTOL_FP32_TC) if queue_dtype == "fp32" else (LOSS_RTOL_BF16, GRAD_RTOL)
    for st in range(int(j["steps"])):
        q, qj = cu(T_(j[f"q{st}"])).requires_grad_(True), cu(T_(j[f"q_jig{st}"])).requires_grad_(True)
        out, out_jig, labels = moco(q, cu(T_(j[f"k{st}"])), q_jig=qj)
        l0, l1 = crit(out), crit(out_jig)
        (l0 + l1).backward()
        for l, x, name in ((l0, q, ""), (l1, qj, "_jig")):
            assert abs(float(l.detach()) - float(j[f"loss{name}{st}"])) <= ltol * float(j[f"loss{name}{st}"])
            assert rel_max(x.grad, T_(j[f"dq{name}{st}"])) <= gtol
        assert moco.index == int(j[f"index_after{st}"]) and labels.shape == (q.shape[0],)
    mem = moco.state_dict()["memory"].cpu()
    ref = T_(j["memory_after"])
    assert torch.equal(mem, ref if queue_dtype == "fp32" else bf16r(ref))


@pytest.mark.parametrize("queue_dtype", ["fp32", "bf16"])
def test_cmcmoco_against_reference_fixture(lib, golden, queue_dtype):
    """`CMCMoCo` (mem_moco.py:91-142): two queues, cross-modal positives, jig heads, gathered keys crossing the end of the
    ring -- every head's loss and gradient, both queues' contents and the shared pointer against the reference's module."""
    import gca_b200
    g = golden("cmc_moco")
    K, T = int(g["K"]), float(g["T"])
    m = gca_b200.CMCMoCo(128, K=K, T=T, queue_dtype=queue_dtype).cuda()
    m.load_state_dict({"memory_1": T_(g["memory_1_before"]), "memory_2": T_(g["memory_2_before"])})
    m.index = int(g["start_index"])
    crit = gca_b200.NCESoftmaxLoss()
    ltol, gtol = (LOSS_RTOL_FP32, GRAD_RTOL_FP32_TC) if queue_dtype == "fp32" else (LOSS_RTOL_BF16, GRAD_RTOL)
    for st in range(int(g["steps"])):
        nh = 4 if f"q{st}_3" in g else 2
        qs = [cu(T_(g[f"q{st}_{i}"])).requires_grad_(True) for i in range(nh)]
        kw = {}
        if nh == 4:
            kw.update(q1_jig=qs[2], q2_jig=qs[3])
        if f"all_k1_{st}" in g:
            kw.update(all_k1=cu(T_(g[f"all_k1_{st}"])), all_k2=cu(T_(g[f"all_k2_{st}"])))
        out = m(qs[0], cu(T_(g[f"k1_{st}"])), qs[1], cu(T_(g[f"k2_{st}"])), **kw)
        heads, labels = out[:-1], out[-1]
        assert len(heads) == nh and labels.dtype == torch.int64 and int(labels.abs().sum()) == 0
        losses = [crit(h) for h in heads]
        sum(losses).backward()
        for i in range(nh):
            assert abs(float(losses[i].detach()) - float(g[f"loss{st}_{i}"])) <= ltol * float(g[f"loss{st}_{i}"]), (st, i)
            assert rel_max(qs[i].grad, T_(g[f"dq{st}_{i}"])) <= gtol, (st, i)
        assert m.index == int(g[f"index_after{st}"])                                    # integer: exact
    sd = m.state_dict()
    for name in ("memory_1", "memory_2"):
        ref = T_(g[name + "_after"])
        assert torch.equal(sd[name].cpu(), ref if queue_dtype == "fp32" else bf16r(ref))  # slot contents: exact


def test_materialized_logits_mode(lib, golden):
    import gca_b200
    g = golden("infonce_small")
    moco = gca_b200.RGBMoCo(128, K=256, T=float(g["T"]), materialize=True).cuda()
    moco.load_state_dict({"memory": T_(g["memory_before"])})
    q = cu(T_(g["q0"])).requires_grad_(True)
    out, labels = moco(q, cu(T_(g["k0"])))
    assert isinstance(out, torch.Tensor) and out.shape == (8, 257)
    loss = gca_b200.NCESoftmaxLoss()(out)
    loss.backward()
    assert abs(float(loss) - float(g["loss0"])) <= LOSS_RTOL_FP32 * float(g["loss0"])
    assert rel_max(q.grad, T_(g["dq0"])) <= 1e-4


# ============================================================================================ K4/K5 graph head
@pytest.mark.parametrize("name", ["graph_c1", "graph_fmap", "graph_odd", "graph_t2"])
def test_graph_core_golden(GF, golden, name):
    g = golden(name)
    alpha, max_hop, temp, sub = float(g["alpha"]), int(g["max_hop"]), float(g["temperature"]), bool(g["sub_sample"])
    x, wq, wk, wg, u, dy = (T_(g[n]) for n in ("x", "wq", "wk", "wg", "u", "dy"))
    B, C, T = x.shape[:3]
    gq = og._project(x, wq, sub, True)
    gk = og._project(x, wk, sub, True)
    sup = F.conv3d(x, wg)
    gq_d, gk_d, sup_d = (cu(t).requires_grad_(True) for t in (gq, gk, sup))
    y, sim, adj, s = GF.graph_core(gq_d, gk_d, sup_d, cu(u), alpha, max_hop, temp)
    assert rel_max(sim, T_(g["sim"])) <= 1e-5
    assert rel_max(adj, T_(g["adj"])) <= 1e-5
    assert rel_max(s, T_(g["s"])) <= 1e-5
    # hop mask is integer work: adj must be exactly zero outside max_hop and non-zero inside
    hop = T_(g["hop"])
    assert torch.equal((adj[0].cpu() == 0), (hop < 0))
    assert rel_max(y.reshape(x.shape), T_(g["y"])) <= 1e-5
    y.backward(cu(dy).reshape(y.shape))
    assert rel_max(gq_d.grad.reshape(g["d_gq"].shape), T_(g["d_gq"])) <= 1e-3
    assert rel_max(gk_d.grad.reshape(g["d_gk"].shape), T_(g["d_gk"])) <= 1e-3
    assert rel_max(sup_d.grad.reshape(g["d_support"].shape), T_(g["d_support"])) <= 1e-4


@pytest.mark.parametrize("name", ["graph_c1", "graph_fmap", "graph_odd"])
def test_graph_module_end_to_end(lib, golden, name, monkeypatch):
    """The drop-in module with the reference's weights and the reference's uniforms reproduces y, dx and dW."""
    import gca_b200
    # the learned 1x1x1 convolutions stay with cuDNN: compare in strict fp32 (TF32 convolutions are cuDNN's default)
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    monkeypatch.setattr(torch.backends.cuda.matmul, "allow_tf32", False)
    g = golden(name)
    x = T_(g["x"])
    m = gca_b200.TemporalGraphAug(x.shape[1], sub_sample=bool(g["sub_sample"]), max_hop=int(g["max_hop"]),
                                  alpah=float(g["alpha"]), temperature=float(g["temperature"]))
    sd = {"gcns.0.conv.weight": T_(g["wg"])}
    pre = "g_q.0.weight" if bool(g["sub_sample"]) else "g_q.weight"
    sd[pre] = T_(g["wq"])
    sd[pre.replace("g_q", "g_k")] = T_(g["wk"])
    m.load_state_dict(sd)
    m = m.cuda()
    u = cu(T_(g["u"]))
    monkeypatch.setattr(torch, "rand", lambda *a, **k: u.clone())       # the one RNG draw of the forward
    xd = cu(x).requires_grad_(True)
    y = m(xd)
    assert y.shape == x.shape and rel_max(y, T_(g["y"])) <= 1e-5
    y.backward(cu(T_(g["dy"])))
    assert rel_max(xd.grad, T_(g["dx"])) <= 1e-3
    conv_q = m.g_q[0] if bool(g["sub_sample"]) else m.g_q
    assert rel_max(conv_q.weight.grad, T_(g["dwq"])) <= 1e-3
    assert rel_max(m.gcns[0].conv.weight.grad, T_(g["dwg"])) <= 1e-3


def test_graph_module_draws_like_reference_rsample(lib):
    """Seeded forward == functional call with u = torch.rand(B,T,T) drawn right after the same seed."""
    import gca_b200
    from gca_b200 import functional as GFm
    torch.manual_seed(0)
    m = gca_b200.TemporalGraphAug(32, sub_sample=False).cuda()
    x = torch.randn(4, 32, 8, 1, 1, device="cuda")
    torch.manual_seed(123)
    y1, info = m(x, return_graph=True)
    torch.manual_seed(123)
    u = torch.rand(4, 8, 8, device="cuda")
    assert torch.equal(info["u"], u)
    y2, _, _, _ = GFm.graph_core(m.g_q(x), m.g_k(x), m.gcns[0].conv(x), u, 0.5, 3, 1.0)
    assert torch.equal(y1, y2.view_as(y1))


@pytest.mark.parametrize("shape,sub", [((3, 32, 8, 28, 28), True), ((2, 24, 16, 14, 14), False), ((2, 8, 32, 6, 6), False),
                                       ((5, 64, 8, 7, 7), False), ((130, 256, 8, 1, 1), False),
                                       ((128, 192, 8, 14, 14), True), ((128, 1024, 8, 1, 1), False)])
def test_graph_core_random_shapes(GF, shape, sub):
    """Large feature maps (split adjacency + grid aggregation path), odd spatial sizes, T up to 32; the last two are
    BASELINE config 3 at full size: the S3D base.5 feature map [128, 192, 8, 14, 14] with sub-sampled projections and
    the embedding-level head [128, 1024, 8, 1, 1]."""
    B, C, T, H, W = shape
    gen = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(*shape, generator=gen)
    # projections scaled so the similarity logits stay O(1) -- saturated softmaxes would only test exp() rounding
    wq = torch.randn(C // 2, C, 1, 1, 1, generator=gen) * (0.3 / (C * H * W) ** 0.5)
    wk = torch.randn(C // 2, C, 1, 1, 1, generator=gen) * (0.3 / C ** 0.5)
    wg = torch.randn(C, C, 1, 1, 1, generator=gen) / C ** 0.5
    u = torch.rand(B, T, T, generator=gen)
    dy = torch.randn(*shape, generator=gen)
    gq = og._project(x, wq, sub, True).reshape(B, C // 2, T, -1)
    gk = og._project(x, wk, sub, True).reshape(B, C // 2, T, -1)
    sup = F.conv3d(x, wg).reshape(B, C, T, -1)
    y_ref, sim, adj, s = og.graph_core(gq, gk, sup, u)
    d_ref = og.graph_core_backward(gq, gk, sup, sim, adj, s, dy.reshape(B, C, T, -1))
    gq_d, gk_d, sup_d = (cu(t).requires_grad_(True) for t in (gq, gk, sup))
    y, sim_d, adj_d, s_d = GF.graph_core(gq_d, gk_d, sup_d, cu(u))
    assert rel_max(sim_d, sim) <= 1e-4 and rel_max(s_d, s) <= 1e-4      # logits of magnitude ~50 over D ~ 3000 terms
    assert rel_max(y, y_ref) <= 1e-4
    y.backward(cu(dy).reshape(y.shape))
    for got, ref in zip((gq_d.grad, gk_d.grad, sup_d.grad), d_ref):
        assert rel_max(got, ref) <= 2e-3


# ============================================================================================ K6 SimSiam D
VARIANTS = [dict(threshold=0.05), dict(topk=3), dict(edge_drop=0.3), dict(symnorm=True),
            dict(threshold=0.02, topk=4, symnorm=True), dict(topk=2, edge_drop=0.25, symnorm=True)]


@pytest.mark.parametrize("shape", [(3, 16, 8, 4, 4), (2, 192, 8, 14, 14), (4, 64, 8, 1, 1), (2, 8, 16, 6, 6)])
@pytest.mark.parametrize("variants", VARIANTS, ids=lambda v: "+".join(sorted(v)))
def test_graph_variants_against_own_restatement(GF, shape, variants):
    """The default-OFF variants of the graph head (GCA_GRAPH_THRESHOLD / TOPK / EDGE_DROP / SYMNORM; no reference counterpart,
    PARITY UNPINNED): forward and backward of gca_graph_fwd_ex / gca_graph_bwd_ex against oracle.graph.graph_core_variants
    (torch autograd), all three kernel paths (per-video, split, large maps).  The kept-edge masks are integer work: exact."""
    B, C, T, H, W = shape
    torch.manual_seed(sum(shape))
    Cq = max(C // 2, 1)
    gq = (torch.randn(B, Cq, T, H * W) * 0.3).requires_grad_(True)
    gk = (torch.randn(B, Cq, T, H * W) * 0.3).requires_grad_(True)
    sup = torch.randn(B, C, T, H * W).requires_grad_(True)
    u = torch.rand(B, T, T)
    dy = torch.randn(B, C, T, H * W)
    y_ref, sim_ref, adj_ref, s_ref = og.graph_core_variants(gq.double(), gk.double(), sup.double(), u.double(), **variants)
    g_ref = torch.autograd.grad(y_ref, [gq, gk, sup], dy.double())
    a, b, c = cu(gq.detach()).requires_grad_(True), cu(gk.detach()).requires_grad_(True), cu(sup.detach()).requires_grad_(True)
    y, sim, adj, s = GF.graph_core(a, b, c, cu(u), variants=variants)
    y.backward(cu(dy))
    # masks: an entry is removed in the kernel exactly where the restatement removes it (away from fp32 ties at the cut)
    kept_ref, kept = adj_ref > 0, adj.cpu() > 0
    near_cut = torch.zeros_like(kept_ref)
    if "threshold" in variants:
        near_cut |= (sim_ref * og.edge_weight_matrix(T, 3, 0.5, torch.float64)[None] - variants["threshold"]).abs() < 1e-6
    if not near_cut.any() and "topk" not in variants:
        assert torch.equal(kept, kept_ref)
    if "topk" in variants and "threshold" not in variants:
        assert int((kept != kept_ref).sum()) <= 2 * B                      # (fp32 near-ties at the k-th entry may swap)
    if torch.equal(kept, kept_ref):
        assert rel_max(s, s_ref) <= 2e-5 and rel_max(y, y_ref) <= 2e-5
        for got, ref in zip((a.grad, b.grad, c.grad), g_ref):
            assert rel_max(got, ref) <= 2e-4


def test_graph_module_variants_compose(lib):
    """Cosine adjacency and the feature mask are compositions around the kernel (module options, default OFF): shapes, the
    seeded masks and autograd through them; the default module draws exactly one torch.rand like the reference."""
    import gca_b200
    torch.manual_seed(0)
    m = gca_b200.TemporalGraphAug(32, sub_sample=False, adjacency="cosine", edge_topk=3, sym_norm=True, feature_mask=0.25).cuda()
    x = torch.randn(4, 32, 8, 2, 2, device="cuda", requires_grad=True)
    torch.manual_seed(7)
    y, info = m(x, return_graph=True)
    torch.manual_seed(7)
    u = torch.rand(4, 8, 8, device="cuda")
    keep = torch.rand(4, 32, device="cuda") >= 0.25
    assert torch.equal(u, info["u"])
    assert torch.equal((y.detach().abs().sum(dim=(2, 3, 4)) > 0), keep)          # masked channels are exactly zero
    assert int((info["adj"] > 0).sum(-1).max()) <= 3                            # top-k rows
    y.sum().backward()
    assert torch.isfinite(x.grad).all() and float(x.grad.abs().sum()) > 0
    # cosine: the logits are bounded by 1 in magnitude -> softmax rows within e^2 of uniform
    assert float(info["sim"].max() / info["sim"].min()) <= np.e ** 2 + 1e-3


def test_negcos_golden(lib, golden):
    import gca_b200
    g = golden("negcos")
    p = cu(T_(g["p"])).requires_grad_(True)
    loss = gca_b200.D()(p, cu(T_(g["z"])))
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) <= 1e-6
    assert rel_max(p.grad, T_(g["dp"])) <= 1e-5


@pytest.mark.parametrize("B,d", [(128, 1024), (7, 2048), (33, 50), (1, 8)])
def test_negcos_random(GF, B, d):
    gen = torch.Generator().manual_seed(B + d)
    p, z = torch.randn(B, d, generator=gen), torch.randn(B, d, generator=gen)
    pd = cu(p).requires_grad_(True)
    loss = GF.neg_cosine(pd, cu(z))
    (loss * 3.0).backward()
    assert abs(float(loss) - float(oracle.neg_cosine(p.double(), z.double()))) <= 1e-6
    assert rel_max(pd.grad, 3.0 * oracle.neg_cosine_grad(p.double(), z.double())) <= 1e-5


# ============================================================================================ K7 retrieval
def test_retrieval_small_golden(GF, golden):
    g = golden("retrieval")
    idx, val = GF.cosine_topk(cu(T_(g["small_queries"])), cu(T_(g["small_gallery"])), 50)
    np.testing.assert_array_equal(idx[:, :10].cpu().numpy(), g["small_top10"])
    hits = oracle.recall_hits(idx.cpu().numpy(), g["small_query_labels"], g["small_gallery_labels"])
    assert [hits[k] for k in oracle.retrieval.KS] == list(g["small_hits"])            # integer recall counts: exact
    assert bool((val[:, :-1] >= val[:, 1:]).all())                                     # sortedness


def test_retrieval_config5_full_size(GF, golden):
    """BASELINE config 5 / SURVEY G3: 3,783 x 13,320 x 512, recall@{1,5,10,20,50} hit counts equal the reference's."""
    g = golden("retrieval")
    rng = np.random.default_rng(0)
    gal = rng.standard_normal((13320, 512)).astype(np.float32)
    qry = rng.standard_normal((3783, 512)).astype(np.float32)
    gl = rng.integers(0, 101, 13320)
    ql = rng.integers(0, 101, 3783)
    idx, val = GF.cosine_topk(cu(T_(qry)), cu(T_(gal)), 50)
    hits = oracle.recall_hits(idx.cpu().numpy(), ql, gl)
    assert [hits[k] for k in oracle.retrieval.KS] == list(g["c5_hits"]) == [28, 151, 349, 666, 1478]
    assert bool((val[:, :-1] >= val[:, 1:]).all())
    assert len(set(idx[0].cpu().tolist())) == 50


@pytest.mark.parametrize("Nq,Ng,d,k", [(200, 1001, 64, 50), (129, 300, 128, 64), (5, 70, 192, 7), (384, 2048, 512, 50)])
def test_retrieval_tensor_core_path_matches_oracle(GF, Nq, Ng, d, k):
    """d % 64 == 0 takes the split-bf16 tcgen05 panel (ragged tile edges, panel rows padded to 16 bytes): indices equal
    the fp32 oracle's stable argsort, similarities to fp32 rounding."""
    rng = np.random.default_rng(Nq + Ng + d)
    qry = rng.standard_normal((Nq, d)).astype(np.float32)
    gal = rng.standard_normal((Ng, d)).astype(np.float32)
    gal[Ng // 2] = gal[3]                                                  # an exact tie: the lower index must come first
    idx, val = GF.cosine_topk(cu(T_(qry)), cu(T_(gal)), k)
    oidx, oval = oracle.cosine_topk(qry, gal, k)
    got, gv = idx.cpu().numpy(), val.cpu().numpy()
    np.testing.assert_allclose(gv, oval, rtol=0, atol=3e-6)
    mism = got != oidx
    if mism.any():                                                         # only near-ties (below fp32 resolution) may swap
        r, c = np.nonzero(mism)
        assert np.all(np.abs(oval[r, c] - gv[r, c]) <= 3e-6) and mism.mean() < 0.002
        for row in np.unique(r):
            assert sorted(got[row].tolist()) == sorted(oidx[row].tolist()) or mism[row, -1]
    three = np.nonzero((got == 3).any(1) & (got == Ng // 2).any(1))[0]
    for row in three:                                                      # tie order: index 3 before its duplicate
        assert list(got[row]).index(3) < list(got[row]).index(Ng // 2)


def test_retrieval_massive_ties_take_the_fallback(GF):
    """More candidates at the threshold than the shared list holds (1500 identical gallery rows): k rounds of arg-max."""
    rng = np.random.default_rng(5)
    gal = rng.standard_normal((1700, 64)).astype(np.float32)
    gal[100:1600] = gal[100]
    qry = np.stack([gal[100], -gal[100], gal[7]]).astype(np.float32)
    idx, val = GF.cosine_topk(cu(T_(qry)), cu(T_(gal)), 20)
    got = idx.cpu().numpy()
    assert got[0].tolist() == list(range(100, 120))                       # all cosine 1: lowest indices first
    assert got[2][0] == 7
    oidx, _ = oracle.cosine_topk(qry, gal, 20)
    assert got[2].tolist() == oidx[2].tolist()
    assert bool((val[:, :-1] >= val[:, 1:]).all())


@pytest.mark.parametrize("Nq,Ng,d,k", [(300, 5000, 128, 50), (131, 4099, 64, 10), (64, 9000, 512, 64)])
def test_retrieval_large_gallery_matches_oracle(GF, Nq, Ng, d, k):
    """Galleries of several thousand rows (many column tiles per row, per-tile maxima as the threshold input): indices equal
    the fp32 oracle's stable argsort (near-ties below fp32 resolution excepted), the exact tie keeps the lower index first."""
    rng = np.random.default_rng(Nq + Ng + d)
    qry = rng.standard_normal((Nq, d)).astype(np.float32)
    gal = rng.standard_normal((Ng, d)).astype(np.float32)
    gal[Ng // 2] = gal[3]
    idx, val = GF.cosine_topk(cu(T_(qry)), cu(T_(gal)), k)
    oidx, oval = oracle.cosine_topk(qry, gal, k)
    got, gv = idx.cpu().numpy(), val.cpu().numpy()
    np.testing.assert_allclose(gv, oval, rtol=0, atol=3e-6)
    mism = got != oidx
    if mism.any():
        r, c = np.nonzero(mism)
        assert np.all(np.abs(oval[r, c] - gv[r, c]) <= 3e-6) and mism.mean() < 0.002
    assert bool((val[:, :-1] >= val[:, 1:]).all())
    three = np.nonzero((got == 3).any(1) & (got == Ng // 2).any(1))[0]
    for row in three:
        assert list(got[row]).index(3) < list(got[row]).index(Ng // 2)


def test_retrieval_large_gallery_massive_ties(GF):
    """2000 identical gallery rows in a 6000-row gallery: the candidate list of a query equal to them overflows and the row
    takes the arg-max rounds (lowest indices first); the other rows come from their candidate lists."""
    rng = np.random.default_rng(11)
    gal = rng.standard_normal((6000, 64)).astype(np.float32)
    gal[1000:3000] = gal[1000]
    qry = np.concatenate([gal[1000:1001], rng.standard_normal((40, 64)).astype(np.float32)])
    idx, val = GF.cosine_topk(cu(T_(qry)), cu(T_(gal)), 20)
    got = idx.cpu().numpy()
    assert got[0].tolist() == list(range(1000, 1020))
    oidx, oval = oracle.cosine_topk(qry, gal, 20)
    np.testing.assert_allclose(val.cpu().numpy(), oval, rtol=0, atol=3e-6)
    assert (got[1:] != oidx[1:]).mean() < 0.01


# ============================================================================================ K13 instance bank
def test_instance_bank_matches_reference_fixture(GF, golden):
    """gca_b200.RGBMem / CMCMem against tests/golden/bank.npz (the reference's own modules): logits, d loss / d x under
    NCESoftmaxLoss (also of the jigsaw head), the bank after the momentum update incl. a duplicated index."""
    import gca_b200
    g = golden("bank")
    n_data, K, T, m = int(g["n_data"]), int(g["K"]), float(g["T"]), float(g["m"])
    bank = gca_b200.RGBMem(128, n_data, K=K, T=T, m=m).cuda()
    bank.load_state_dict({"memory": T_(g["rgb_memory_before"])})
    crit = gca_b200.NCESoftmaxLoss()
    for st in range(2):
        x = cu(T_(g["rgb%d_x" % st])).requires_grad_(True)
        y, idx = cu(T_(g["rgb%d_y" % st])), cu(T_(g["rgb%d_idx" % st]))
        if st == 1:
            xj = cu(T_(g["rgb1_x_jig"])).requires_grad_(True)
            logits, logits_jig, labels = bank(x, y, xj, cu(T_(g["rgb1_all_x"])), cu(T_(g["rgb1_all_y"])), idx=idx)
            (crit(logits) + crit(logits_jig)).backward()
            assert rel_max(logits_jig, T_(g["rgb1_logits_jig"])) <= 2e-6 and rel_max(xj.grad, T_(g["rgb1_dx_jig"])) <= 1e-5
        else:
            logits, labels = bank(x, y, idx=idx)
            crit(logits).backward()
        assert labels.dtype == torch.long and int(labels.abs().sum()) == 0
        assert rel_max(logits, T_(g["rgb%d_logits" % st])) <= 2e-6
        assert rel_max(x.grad, T_(g["rgb%d_dx" % st])) <= 1e-5
        after = T_(g["rgb%d_memory_after" % st])
        assert float((bank.memory.cpu() - after).abs().max()) <= 2e-7
    two = gca_b200.CMCMem(128, n_data, K=K, T=T, m=m).cuda()
    two.load_state_dict({"memory_1": T_(g["cmc_memory_1_before"]), "memory_2": T_(g["cmc_memory_2_before"])})
    x1, x2 = cu(T_(g["cmc_x1"])).requires_grad_(True), cu(T_(g["cmc_x2"])).requires_grad_(True)
    l1, l2, labels = two(x1, x2, cu(T_(g["cmc_y"])), idx=cu(T_(g["cmc_idx"])))
    (crit(l1) + crit(l2)).backward()
    assert rel_max(l1, T_(g["cmc_logits1"])) <= 2e-6 and rel_max(l2, T_(g["cmc_logits2"])) <= 2e-6
    assert rel_max(x1.grad, T_(g["cmc_dx1"])) <= 1e-5 and rel_max(x2.grad, T_(g["cmc_dx2"])) <= 1e-5
    assert float((two.memory_1.cpu() - T_(g["cmc_memory_1_after"])).abs().max()) <= 2e-7
    assert float((two.memory_2.cpu() - T_(g["cmc_memory_2_after"])).abs().max()) <= 2e-7


@pytest.mark.parametrize("B,K,d,n_data", [(7, 33, 64, 500), (64, 4096, 128, 20000), (16, 1000, 320, 3000), (256, 16384, 128, 100000)])
def test_instance_bank_kernels_match_oracle(GF, B, K, d, n_data):
    """gca_bank_logits / gca_bank_dx / gca_bank_update at ragged and at full size (256 x 16385 sampled rows) against
    oracle/bank.py: logits and gradients to fp32 rounding, the drawn indices stay inside the bank, the update with
    duplicated indices keeps the LAST occurrence and leaves every other row untouched (bit-exact)."""
    from oracle import bank as ob
    import gca_b200
    gen = torch.Generator().manual_seed(B + K + d)
    mem = unit_rows(n_data, d, gen)
    x = unit_rows(B, d, gen)
    idx = torch.randint(0, n_data, (B, K + 1), generator=gen)
    xg = cu(x).requires_grad_(True)
    logits = GF.bank_logits(xg, cu(mem), cu(idx), 0.07)
    w = torch.randn(B, K + 1, generator=gen)
    (logits * cu(w)).sum().backward()
    # (the restatement materialises the [rows, K+1, d] gather in fp64: 32 rows at a time keep it at half a gigabyte)
    md = mem.double()
    ref = torch.cat([ob.bank_logits(x[i:i + 32].double(), md, idx[i:i + 32], 0.07) for i in range(0, B, 32)])
    assert rel_max(logits, ref) <= 2e-6
    ref_dx = torch.cat([ob.bank_grad_x(w[i:i + 32].double(), md, idx[i:i + 32], 0.07) for i in range(0, B, 32)])
    assert rel_max(xg.grad, ref_dx) <= 2e-5
    # update: N rows, some indices twice
    N = min(2 * B, 300)
    y = torch.randint(0, n_data, (N,), generator=gen)
    y[N // 2] = y[1]
    y[N - 1] = y[0]
    feats = unit_rows(N, d, gen)
    dmem = cu(mem).clone()
    GF.bank_update_(dmem, cu(feats), cu(y), 0.5)
    expect = ob.bank_update(mem.clone(), feats, y, 0.5)
    got = dmem.cpu()
    touched = torch.zeros(n_data, dtype=torch.bool)
    touched[y] = True
    assert torch.equal(got[~touched], mem[~touched])                      # untouched rows: bit-exact
    assert float((got[touched] - expect[touched]).abs().max()) <= 2e-7
    # the module draws in-range indices with the positive in column 0
    bank = gca_b200.RGBMem(d, n_data, K=K).cuda()
    yy = cu(torch.randint(0, n_data, (B,), generator=gen))
    with torch.no_grad():
        lg, labels = bank(cu(x), yy)
    assert lg.shape == (B, K + 1) and bool(torch.isfinite(lg).all())


# ============================================================================================ full-size properties
def test_headline_shape_full_size(GF):
    """BASELINE metric shape (B=256, K=65536, d=128): both modes against the fp64 oracle, pointer wrap over a lap."""
    import gca_b200
    gen = torch.Generator().manual_seed(1)
    mem32 = unit_rows(65536, 128, gen)
    q, k = unit_rows(256, 128, gen), unit_rows(256, 128, gen)
    check_infonce(GF, q, k, mem32, 0.07, "ffma", LOSS_RTOL_FP32, 1e-3)
    check_infonce(GF, q, k, mem32.to(torch.bfloat16), 0.07, "tcgen05", LOSS_RTOL_BF16, GRAD_RTOL, ref_mem=mem32)
    moco = gca_b200.RGBMoCo(128, K=65536, queue_dtype="bf16").cuda()
    moco.index = 65536 - 256 * 2
    ref_mem = moco.memory.cpu().clone()
    idx = moco.index
    for _ in range(4):                                                    # crosses the end of the ring
        kk = unit_rows(256, 128, gen)
        moco(cu(unit_rows(256, 128, gen)), cu(kk))
        idx = oracle.enqueue(ref_mem, kk, idx)
    assert moco.index == idx == 512 and torch.equal(moco.memory.cpu(), ref_mem)


# ============================================================================================ captured step (CUDA graph)
@pytest.mark.parametrize("queue_dtype,B,N,K", [("bf16", 64, 64, 1024), ("fp32", 32, 48, 256), ("bf16", 256, 256, 4096)])
def test_graphed_step_matches_oracle_over_a_lap(lib, queue_dtype, B, N, K):
    """gca_moco_step replayed from a CUDA graph: loss, gradient, top-k hits, queue contents and the device-resident ring
    pointer after enough steps to wrap the ring (integer state bit-exact against the oracle)."""
    import gca_b200
    from gca_b200.graphed import GraphedMoCoStep
    gen = torch.Generator().manual_seed(B + N + K)
    moco = gca_b200.RGBMoCo(128, K=K, T=0.07, queue_dtype=queue_dtype).cuda()
    moco.index = K - N - 5                                   # the second step wraps
    ref_mem = moco.memory.cpu().clone()
    idx = moco.index
    step = GraphedMoCoStep(moco, B, N).capture()
    # bf16 (tcgen05): prep + streaming kernel + finalize (with the enqueue riding in it); fp32 (ffma family): no prep launch
    assert step.launches_per_step == (3 if queue_dtype == "bf16" else (4 if moco.memory.shape[1] == 128 else 2))
    for it in range(K // N + 3):
        q, k, all_k = unit_rows(B, 128, gen), unit_rows(B, 128, gen), unit_rows(N, 128, gen)
        prev_mem = ref_mem.clone()
        loss = step.step(cu(q), cu(k), cu(all_k))
        rq = tcq(q, moco.T) if queue_dtype == "bf16" else q
        o = oracle.infonce_step(rq.double(), k.double(), ref_mem.double(), 0, 0.07)
        idx = oracle.enqueue(ref_mem, all_k, idx)
        tol = LOSS_RTOL_BF16 if queue_dtype == "bf16" else LOSS_RTOL_FP32
        assert abs(float(loss) - float(o["loss"])) <= tol * abs(float(o["loss"]))
        assert rel_max(step.dq, o["dq"]) <= GRAD_RTOL
        hits = step.hits.cpu().tolist()
        assert hits[0] <= hits[1] <= B
        # top-1 / top-5 hit counts (what `accuracy` reports, train...:428) against the oracle's ranks, up to the rows whose
        # nearest negative sits inside the accumulation noise of the positive
        # (the kernels take the positive from the unrounded fp32 q.k, the negatives from the bf16-rounded queries)
        neg = oracle.logits_full(rq.double(), k.double(), prev_mem.double(), 0.07)[:, 1:]
        pos = ((q.double() * k.double()).sum(1) / 0.07)[:, None]
        rank_ref = (neg > pos).sum(1)
        clear = (neg - pos).abs().min(1).values > 3e-4
        unsure = int((~clear).sum())
        assert abs(hits[0] - int((rank_ref < 1).sum())) <= unsure and abs(hits[1] - int((rank_ref < 5).sum())) <= unsure
        assert torch.equal(step.rank.cpu().long()[clear], rank_ref[clear])
        assert moco.index == idx
    torch.cuda.synchronize()
    assert int(step.state[0]) == idx and int(step.state[1]) == 0           # device pointer == host pointer == oracle
    assert torch.equal(moco.memory.cpu(), ref_mem)                          # slot contents: exact


@pytest.mark.parametrize("B,N,K", [(64, 64, 1024), (256, 256, 65536)])
def test_launch_plan_equals_graph_replay(lib, B, N, K):
    """gca_plan_*: the step re-issued from a recorded launch plan (three launches, programmatic dependent launch between
    them and across consecutive steps) gives bit-identical loss, gradient, hit counts, queue and ring pointer to the
    CUDA-graph replay of the same step, back to back over a wrap of the ring; launch bookkeeping agrees."""
    import gca_b200
    from gca_b200.graphed import GraphedMoCoStep
    gen = torch.Generator().manual_seed(7 * B + K)
    mocos = [gca_b200.RGBMoCo(128, K=K, T=0.07, queue_dtype="bf16").cuda() for _ in range(2)]
    mocos[1].load_state_dict(mocos[0].state_dict())
    for m in mocos:
        m.index = K - 2 * N - 3
    steps = [GraphedMoCoStep(m, B, N).capture() for m in mocos]
    assert steps[0].plan is not None and steps[0].plan.launches == steps[0].launches_per_step == 3
    steps[1].prefer_graph = True
    n0 = lib.gca_launch_count()
    nsteps = 6
    for it in range(nsteps):                                              # no synchronisation between the steps of a variant
        q, k, all_k = cu(unit_rows(B, 128, gen)), cu(unit_rows(B, 128, gen)), cu(unit_rows(N, 128, gen))
        for s in steps:
            s.step(q, k, all_k)
        if it in (1, nsteps - 1):
            torch.cuda.synchronize()
            assert torch.equal(steps[0].outputs, steps[1].outputs)        # loss | hits | dq
            assert torch.equal(steps[0].rank, steps[1].rank) and torch.equal(steps[0].lse, steps[1].lse)
    torch.cuda.synchronize()
    assert int(lib.gca_launch_count() - n0) == 3 * nsteps                 # the plan counts its launches, graph replays do not
    assert torch.equal(steps[0].state, steps[1].state) and mocos[0].index == mocos[1].index
    assert torch.equal(mocos[0].memory, mocos[1].memory)


def test_launch_plan_of_the_projection_step(lib, GF):
    """gca_moco_step_proj is recordable too (same launches as the plain step): a plan over static buffers with the
    device-resident ring pointer, run three times, equals three direct calls -- loss, dz, normalised keys, queue, pointer."""
    import ctypes
    from gca_b200 import _lib
    from gca_b200._lib import ptr
    gen = torch.Generator().manual_seed(19)
    B, K, T = 96, 2048, 0.07
    mem0 = cu(unit_rows(K, 128, gen)).to(torch.bfloat16)
    zs = [(cu(torch.randn(B, 128, generator=gen) * 2.0), cu(torch.randn(B, 128, generator=gen) * 0.3)) for _ in range(3)]
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    res = []
    for planned in (False, True):
        mem = mem0.clone()
        zq, zk = torch.empty(B, 128, device="cuda"), torch.empty(B, 128, device="cuda")
        state = torch.tensor([K - B - 7, 0], dtype=torch.int64, device="cuda")          # the second step wraps
        loss = torch.zeros(1, device="cuda"); rows = torch.zeros(B, device="cuda"); lse = torch.zeros(B, device="cuda")
        pos = torch.zeros(B, device="cuda"); rank = torch.zeros(B, dtype=torch.int32, device="cuda")
        hits = torch.zeros(2, dtype=torch.int32, device="cuda"); dz = torch.zeros(B, 128, device="cuda"); kh = torch.zeros(B, 128, device="cuda")
        ws = torch.zeros(GF.infonce_workspace_bytes(B, K, 128, 1, "tcgen05"), dtype=torch.uint8, device="cuda")

        def issue(stream):
            _lib.call("gca_moco_step_proj", ptr(zq), ptr(zk), ptr(mem), 1, B, K, 128, 1.0 / T, 2, None, B, 0, ptr(state),
                      ptr(loss), ptr(rows), ptr(lse), ptr(pos), ptr(rank), ptr(hits), ptr(dz), ptr(kh), ptr(ws), ws.numel(), stream)
        plan = _lib.LaunchPlan.record(lambda: issue(None), mem.device) if planned else None
        outs = []
        for a, b in zs:
            zq.copy_(a); zk.copy_(b)
            if planned:
                plan.run(st)
            else:
                issue(st)
            outs.append((loss.clone(), dz.clone(), kh.clone(), rank.clone()))
        torch.cuda.synchronize()
        res.append((outs, mem, state.clone()))
        if planned:
            assert plan.launches == 3
    for (l0, d0, k0, r0), (l1, d1, k1, r1) in zip(res[0][0], res[1][0]):
        assert torch.equal(l0, l1) and torch.equal(d0, d1) and torch.equal(k0, k1) and torch.equal(r0, r1)
    assert torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2])
    assert int(res[1][2][0]) == (K - B - 7 + 3 * B) % K


def test_launch_plan_rejects_unrecordable_calls(lib, GF):
    """Only the tcgen05-family step entry points are recordable: a plan around an fp32-queue step is refused (the call
    itself has then run normally), and begin / end / run misuse returns error codes."""
    import ctypes
    import gca_b200
    from gca_b200 import _lib
    from gca_b200.graphed import GraphedMoCoStep
    assert lib.gca_plan_end(ctypes.byref(ctypes.c_void_p())) == -1        # not recording
    moco = gca_b200.RGBMoCo(64, K=256, T=0.07, queue_dtype="fp32").cuda()
    step = GraphedMoCoStep(moco, 32, 32).capture()
    assert step.plan is None                                              # d = 64: graph replays only
    with pytest.raises(_lib.GcaError, match="recordable"):
        _lib.LaunchPlan.record(lambda: step._enqueue_work(None), moco.memory.device)
    torch.cuda.synchronize()
    assert lib.gca_plan_begin() == 0 and lib.gca_plan_begin() == -1       # already recording
    h = ctypes.c_void_p()
    assert lib.gca_plan_end(ctypes.byref(h)) == -2 and not h.value        # nothing recorded
    assert lib.gca_plan_run(None, None) == -1


@pytest.mark.parametrize("B,K,with_all_k", [(64, 2048, False), (200, 4096, True), (256, 65536, False)])
def test_projection_tail_fusion_matches_normalize_then_head(lib, B, K, with_all_k):
    """gca_moco_step_proj / RGBMoCo.forward_from_projections: Normalize(2) of both projections inside the kernels ==
    F.normalize followed by the ordinary head, in loss, in the gradient w.r.t. the UN-normalised zq (through the
    normalisation) and in the enqueued keys."""
    import gca_b200
    gen = torch.Generator().manual_seed(B + K)
    T = 0.07
    moco = gca_b200.RGBMoCo(128, K=K, T=T, queue_dtype="bf16").cuda()
    moco.index = K - B // 2                                             # the enqueue wraps
    mem0 = moco.memory.float().cpu().clone()
    zq = (torch.randn(B, 128, generator=gen) * 3.0)
    zk = (torch.randn(B, 128, generator=gen) * 0.2)
    all_k = unit_rows(B + 24, 128, gen) if with_all_k else None
    zq_g = cu(zq).requires_grad_(True)
    out, labels, k_hat = moco.forward_from_projections(zq_g, cu(zk), all_k=None if all_k is None else cu(all_k))
    loss = gca_b200.NCESoftmaxLoss()(out)
    loss.backward()
    torch.cuda.synchronize()
    # reference arithmetic in fp64 on the CPU: normalise, then the head on the bf16-rounded query (the kernel's arithmetic)
    zq64 = zq.double().requires_grad_(True)
    qh = zq64 / zq64.norm(dim=1, keepdim=True).clamp_min(1e-12)
    kh = zk.double() / zk.double().norm(dim=1, keepdim=True).clamp_min(1e-12)
    assert rel_max(k_hat, kh) <= 2e-6
    pos = (qh * kh).sum(1) / T
    neg = (qh @ mem0.double().t()) / T
    loss_ref = (torch.logsumexp(torch.cat([pos[:, None], neg], 1), 1) - pos).mean()
    loss_ref.backward()
    assert abs(float(loss.detach()) - float(loss_ref.detach())) <= LOSS_RTOL_BF16 * abs(float(loss_ref.detach()))
    assert rel_max(zq_g.grad, zq64.grad) <= GRAD_RTOL, rel_max(zq_g.grad, zq64.grad)
    # the gradient is orthogonal to zq (scale invariance of the normalised head)
    assert float((zq_g.grad.double().cpu() * zq.double()).sum(1).abs().max()) <= 1e-3 * float(zq_g.grad.abs().max()) * float(zq.norm(dim=1).max())
    # queue: exactly the RNE-bf16 of the keys the call reports (or of all_k), at the wrapped slots
    keys = k_hat.cpu() if all_k is None else all_k
    ref_mem = mem0.clone()
    idx = oracle.enqueue(ref_mem, keys.to(torch.bfloat16).float(), K - B // 2)
    assert moco.index == idx and torch.equal(moco.memory.float().cpu(), ref_mem)
    assert labels.shape == (B,) and int(labels.sum()) == 0


def test_momentum_update_against_reference_fixture(lib, golden):
    """tests/golden/ema.npz: parameters before / after two calls of the reference's own Trainer._momentum_update
    (m = 0.999, then 0.5; oracle/gen_golden_dist.py)."""
    from gca_b200.ema import MomentumUpdater
    g = golden("ema")
    n = int(g["n"])

    class Bag(torch.nn.Module):
        def __init__(self, key):
            super().__init__()
            self.ps = torch.nn.ParameterList([torch.nn.Parameter(cu(T_(g["%s%d" % (key, i)]).clone())) for i in range(n)])
    model, ema = Bag("src"), Bag("before")
    up = MomentumUpdater(model, ema)
    for m in g["ms"]:
        up.step(float(m))
    torch.cuda.synchronize()
    for i, p in enumerate(ema.parameters()):
        assert rel_max(p, T_(g["after%d" % i])) <= 1e-6, i


def test_projection_tail_fusion_against_reference_fixture(lib, golden):
    """The fixture was produced by the reference's own Normalize + RGBMoCo + NCESoftmaxLoss + autograd
    (oracle/gen_golden_proj.py): fp32 loss / gradient / enqueued rows; the kernels run the bf16-queue mode, so the
    BASELINE bf16 tolerances apply (loss 2e-3, gradient 1e-2) and the enqueued rows are compared after bf16 rounding."""
    import gca_b200
    g = golden("proj_tail")
    T = float(g["T"])
    B, K = g["zq"].shape[0], g["memory_before"].shape[0]
    moco = gca_b200.RGBMoCo(128, K=K, T=T, queue_dtype="bf16").cuda()
    moco.load_state_dict({"memory": T_(g["memory_before"])})
    zq = cu(T_(g["zq"])).requires_grad_(True)
    out, labels, k_hat = moco.forward_from_projections(zq, cu(T_(g["zk"])))
    loss = gca_b200.NCESoftmaxLoss()(out)
    loss.backward()
    torch.cuda.synchronize()
    assert abs(float(loss.detach()) - float(g["loss"])) <= LOSS_RTOL_BF16 * float(g["loss"])
    assert rel_max(zq.grad, T_(g["dz"])) <= GRAD_RTOL
    assert moco.index == int(g["index_after"])
    assert rel_max(k_hat, T_(g["enqueued_rows"])) <= 2e-6
    got = moco.memory[:B].float().cpu()
    ref = T_(g["enqueued_rows"]).to(torch.bfloat16).float()
    assert float((got != ref).float().mean()) <= 2e-3          # only values that sit on a bf16 rounding boundary may differ
    assert float((got - ref).abs().max()) <= 2.0 ** -8          # ... and then by one bf16 ulp of a unit-row entry at most


@pytest.mark.parametrize("zero_copy_out,zero_copy_in", [(True, False), (False, False), (True, True)])
def test_graphed_step_host_io_equals_device_step(lib, zero_copy_out, zero_copy_in):
    """step_host_io(): the H2D copy of the pinned inputs, the step and the D2H copy of loss | hits | dq as ONE graph; same
    bits as step() on device-resident inputs, queue and pointer advance identically."""
    import gca_b200
    from gca_b200.graphed import GraphedMoCoStep
    gen = torch.Generator().manual_seed(77)
    B, K = 64, 2048
    mocos = []
    for _ in range(2):
        torch.manual_seed(4)
        mocos.append(gca_b200.RGBMoCo(128, K=K, T=0.07, queue_dtype="bf16").cuda())
    a = GraphedMoCoStep(mocos[0], B, B).capture()
    b = GraphedMoCoStep(mocos[1], B, B)
    host_in = torch.empty(3 * B, 128).pin_memory()
    host_out = torch.empty(b.outputs.shape).pin_memory()
    b.capture_host_io(host_in, host_out, zero_copy_out=zero_copy_out, zero_copy_in=zero_copy_in)
    with pytest.raises(ValueError):
        b.capture_host_io(torch.empty(3 * B, 128), host_out)               # pageable memory is refused
    for it in range(3):
        pk = torch.cat([unit_rows(B, 128, gen), unit_rows(B, 128, gen), unit_rows(B, 128, gen)])
        host_in.copy_(pk)
        a.step(cu(pk[:B]), cu(pk[B:2 * B]), cu(pk[2 * B:]))
        # wait=True: with zero-copy outputs the host polls the step's completion word (the pad word [3] of the result block)
        # instead of synchronising the stream -- the results must be complete the moment it returns
        b.step_host_io(wait=True)
        got = host_out.clone()
        if zero_copy_out:
            assert int(host_out[3:4].view(torch.int32)) == it + 1                            # steps completed on this workspace
            host_in[:2 * B].zero_()                                                          # inputs are consumed: free to overwrite
        torch.cuda.synchronize()
        ref = a.outputs.cpu()
        assert torch.equal(got[:3], ref[:3]) and torch.equal(got[4:], ref[4:])
        assert float(got[0]) == float(a.loss)
        assert torch.equal(mocos[0].memory, mocos[1].memory) and mocos[0].index == mocos[1].index == (it + 1) * B


# ============================================================================================ EMA (momentum encoder update)
@pytest.mark.parametrize("layout", ["contiguous", "channels_last_3d"])
def test_momentum_update_matches_reference_loop(lib, layout):
    """One multi-tensor launch == the per-parameter mul_/add_ loop of Trainer._momentum_update (train...:176-180)."""
    from gca_b200.ema import MomentumUpdater
    torch.manual_seed(0)
    def make():
        return torch.nn.Sequential(torch.nn.Conv3d(3, 17, 3), torch.nn.BatchNorm3d(17), torch.nn.Linear(17, 33),
                                   torch.nn.Linear(33, 5, bias=False), torch.nn.Conv3d(17, 64, (1, 3, 3))).cuda()
    model, ema = make(), make()
    if layout == "channels_last_3d":                                    # permuted-dense conv weights, same on both sides
        model, ema = model.to(memory_format=torch.channels_last_3d), ema.to(memory_format=torch.channels_last_3d)
    ref = [p.detach().clone() for p in ema.parameters()]
    up = MomentumUpdater(model, ema)
    assert up.numel == sum(p.numel() for p in model.parameters())
    for m in (0.999, 0.5, 0.0, 1.0):
        up.step(m)
        for r, p in zip(ref, model.parameters()):
            r.mul_(m).add_(p.detach(), alpha=1 - m)                     # the reference's arithmetic
        torch.cuda.synchronize()
        for r, e in zip(ref, ema.parameters()):
            assert rel_max(e, r) <= 1e-6
    # m = 1 leaves the EMA weights bit-identical, m = 0 copies the online weights exactly
    before = [p.detach().clone() for p in ema.parameters()]
    up.step(1.0)
    assert all(torch.equal(a, b) for a, b in zip(before, ema.parameters()))
    up.step(0.0)
    assert all(torch.equal(a, b) for a, b in zip(model.parameters(), ema.parameters()))
    # the drop-in function keeps its descriptor table on model_ema and reuses it
    from gca_b200.ema import momentum_update
    ref2 = [p.detach().clone() for p in ema.parameters()]
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.25)
    for _ in range(2):
        momentum_update(model, ema, 0.9)
        for r, p in zip(ref2, model.parameters()):
            r.mul_(0.9).add_(p.detach(), alpha=0.1)
    torch.cuda.synchronize()
    assert all(rel_max(e, r) <= 1e-6 for r, e in zip(ref2, ema.parameters()))
    from gca_b200.ema import _UPDATERS
    assert _UPDATERS[ema][1].numel == up.numel
    with pytest.raises(ValueError):                                     # mismatched layouts are refused, not mis-indexed
        MomentumUpdater(make(), make().to(memory_format=torch.channels_last_3d))


@pytest.mark.parametrize("name", ["proj", "pred"])
def test_simsiam_mlp_against_reference_fixture(lib, golden, name):
    """gca_b200.ProjectionMLP / PredictionMLP (library GEMM + fused BatchNorm1d/ReLU launches) against the outputs, gradients
    and running statistics of the reference's own modules (lib/modeling/project_head.py:36-76; oracle/gen_golden_mlp.py)."""
    import gca_b200
    g = golden("simsiam_mlp")
    sd = {k[len(name) + 8:]: T_(v) for k, v in g.items() if k.startswith(name + ".before.")}
    x = T_(g[name + ".x"])
    hid = sd["l1.0.weight"].shape[0]
    m = (gca_b200.ProjectionMLP(x.shape[1], hid, sd["l3.0.weight"].shape[0]) if name == "proj"
         else gca_b200.PredictionMLP(x.shape[1], hid, sd["l2.weight"].shape[0]))
    assert set(m.state_dict().keys()) == set(sd.keys())                     # same checkpoint layout as upstream
    m.load_state_dict(sd)
    m = m.cuda().train()
    xg = x.cuda().requires_grad_(True)
    y = m(xg)
    (y * T_(g[name + ".w"]).cuda()).sum().backward()
    assert rel_max(y, T_(g[name + ".y"])) <= 2e-5
    assert rel_max(xg.grad, T_(g[name + ".dx"])) <= 1e-4
    for k, p_ in m.named_parameters():
        ref = T_(g[name + ".grad." + k])
        # (a bias in front of a BatchNorm has a mathematically zero gradient: only fp32 noise, compared absolutely)
        tol = 1e-4 if k.endswith(".0.bias") else 1e-4 * max(float(ref.abs().max()), 1e-3)
        assert float((p_.grad.cpu() - ref).abs().max()) <= tol, k
    for k, v in m.state_dict().items():
        ref = T_(g[name + ".after." + k])
        if k.endswith("num_batches_tracked"):
            assert int(v) == int(ref)
        elif "running" in k:
            assert float((v.cpu() - ref).abs().max()) <= 1e-5, k
    # eval mode: the running statistics normalise and nothing moves (same as the torch modules on the same state)
    m.eval()
    rm = m.l1[1].running_mean.clone()
    with torch.no_grad():
        ye = m(x.cuda())
    ref_m = torch.nn.Sequential()
    import torch.nn as nn
    l1 = nn.Sequential(nn.Linear(x.shape[1], hid), nn.BatchNorm1d(hid), nn.ReLU()).cuda().eval()
    l1.load_state_dict({k[3:]: v for k, v in m.state_dict().items() if k.startswith("l1.")})
    with torch.no_grad():
        h_ref = l1(x.cuda())
        h = gca_b200.functional.bn1d(torch.nn.functional.linear(x.cuda(), m.l1[0].weight, m.l1[0].bias), m.l1[1], relu=True)
    assert rel_max(h, h_ref) <= 1e-5 and torch.equal(rm, m.l1[1].running_mean) and ye.shape == y.shape


@pytest.mark.parametrize("B,C,relu", [(128, 2048, True), (256, 512, False), (300, 96, True), (2, 33, True), (1024, 64, True)])
def test_bn1d_random_shapes(GF, B, C, relu):
    """gca_bn1d_fwd / gca_bn1d_bwd against torch's own BatchNorm1d (+ ReLU) in fp64 on the CPU: cached (B <= 256) and
    re-reading (B > 256) variants, column tails (C % 32 != 0)."""
    import torch.nn as nn
    torch.manual_seed(B * 7 + C)
    bn = nn.BatchNorm1d(C)
    bn.weight.data.uniform_(0.5, 1.5); bn.bias.data.uniform_(-0.3, 0.3)
    ref = nn.BatchNorm1d(C).double()
    ref.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in bn.state_dict().items()})
    x = torch.randn(B, C) * 1.7 + 0.4
    w = torch.randn(B, C)
    xr = x.double().requires_grad_(True)
    yr = ref(xr)
    yr = torch.relu(yr) if relu else yr
    (yr * w.double()).sum().backward()
    bn = bn.cuda()
    xg = x.cuda().requires_grad_(True)
    y = GF.bn1d(xg, bn, relu=relu)
    (y * w.cuda()).sum().backward()
    assert rel_max(y, yr.detach()) <= 1e-5
    assert rel_max(xg.grad, xr.grad) <= 2e-4
    assert rel_max(bn.weight.grad, ref.weight.grad) <= 1e-4 and rel_max(bn.bias.grad, ref.bias.grad) <= 1e-4
    assert rel_max(bn.running_var, ref.running_var) <= 1e-5 and float((bn.running_mean.cpu() - ref.running_mean).abs().max()) <= 1e-6
    assert int(bn.num_batches_tracked) == 1


def test_graphed_sharded_step_matches_eager_path():
    """GraphedShardedStep (NCCL collectives + kernels in one CUDA graph) == the eager ShardedRGBMoCo path, bit for bit.
    World size 1 here (the suite sees one GPU); `torchrun --nproc-per-node N tests/sharded_graph_worker.py` runs the
    same check over N GPUs.  Separate process: see the worker's note on communicator teardown."""
    import subprocess
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "sharded_graph_worker.py")],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and "SHARDED_GRAPH_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])


def test_peer_key_exchange_single_rank():
    """gca_keys_exchange against NCCL all-gather, eager and graph-replayed (world size 1 inside the suite: the mailbox
    protocol, parity double-buffering and the device-resident step counter; `torchrun --nproc-per-node N
    tests/peer_exchange_worker.py` runs the same check across N GPUs over NVLink)."""
    import subprocess
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "peer_exchange_worker.py")],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and "PEER_EXCHANGE_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])


def _torchrun(worker, nproc, port, timeout=600):
    import subprocess
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(os.path.dirname(os.path.abspath(__file__)), worker)]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (NCCL)")
def test_shuffle_bn_over_nccl_against_reference_fixture():
    """`gca_b200.dist.ShuffleBN` on two GPUs over NCCL against the fixture produced by the reference's `_shuffle_bn`."""
    r = _torchrun("shuffle_bn_worker.py", 2, 29731)
    assert r.returncode == 0 and "SHUFFLE_BN_NCCL_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (NCCL + NVLink peer memory)")
@pytest.mark.parametrize("worker,token", [("sharded_graph_worker.py", "SHARDED_GRAPH_OK"), ("peer_exchange_worker.py", "PEER_EXCHANGE_OK")])
def test_multi_gpu_workers(worker, token):
    """The K-sharded step (eager == graph == ORACLE) and the peer-memory key exchange on every visible GPU (2, 4 or 8)."""
    n = 8 if torch.cuda.device_count() >= 8 else 4 if torch.cuda.device_count() >= 4 else 2
    r = _torchrun(worker, n, 29741 + len(worker))
    assert r.returncode == 0 and token in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
