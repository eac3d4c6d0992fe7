#!/usr/bin/env python
"""Generate tests/golden/*.npz by RUNNING THE REFERENCE's own modules on seeded CPU inputs, and
check every oracle restatement against them while doing so.  TEST INFRASTRUCTURE ONLY.

Runs only in the build container (needs /root/reference, which does not exist on the GPU box);
the fixtures it writes are committed.  Two non-invasive monkeypatches, no reference file is edited
(SURVEY.md R3, R9):
  * torch.Tensor.cuda -> identity           (`.cuda()` is hard-coded at mem_moco.py:25,78 ...)
  * TemporalGraphAug.reset_parameter = reset_parameters   (ctor typo, temporal_graph.py:117,124)

usage:  python oracle/gen_golden.py [--ref /root/reference] [--out tests/golden]
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.dont_write_bytecode = True


def load_reference(ref):
    sys.path.insert(0, ref)
    torch.Tensor.cuda = lambda self, *a, **k: self
    from lib.memory.mem_moco import RGBMoCo
    from lib.memory.criterion import NCESoftmaxLoss, D
    from lib.ops.module_wrappers import temporal_graph as tg
    tg.TemporalGraphAug.reset_parameter = tg.TemporalGraphAug.reset_parameters
    return RGBMoCo, NCESoftmaxLoss, D, tg


def ref_accuracy(output, topk=(1, 5)):
    """lib/evaluation/metric.py:44-67 imported would raise on `.view`; run its
    arithmetic through the reference's own topk call with the `.reshape` fix."""
    maxk = max(topk)
    B = output.shape[0]
    _, pred = output.topk(maxk, 1, True, True)
    correct = pred.t().eq(torch.zeros(B, dtype=torch.long).view(1, -1).expand_as(pred.t()))
    return [float(correct[:k].reshape(-1).float().sum() * (100.0 / B)) for k in topk]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(HERE), "tests", "golden"))
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    RGBMoCo, NCESoftmaxLoss, D, tg = load_reference(args.ref)
    import oracle
    from oracle import graph as og

    torch.set_num_threads(1)   # deterministic reduction order for the fixtures

    # ------------------------------------------------------------------ InfoNCE / queue cases
    def run_moco_case(seed, B, K, d, T, n_all=None, start_index=0, steps=1):
        torch.manual_seed(seed)
        moco = RGBMoCo(d, K=K, T=T)
        moco.index = start_index
        crit = NCESoftmaxLoss()
        mem0 = moco.memory.clone()
        rec = {"memory_before": mem0.numpy().copy(), "start_index": np.int64(start_index),
               "T": np.float64(T)}
        my_mem, my_idx = mem0.clone(), start_index
        for st in range(steps):
            q = F.normalize(torch.randn(B, d)).requires_grad_(True)
            k = F.normalize(torch.randn(B, d))
            all_k = F.normalize(torch.randn(n_all, d)) if n_all else None
            out, labels = moco(q, k, all_k=all_k)
            loss = crit(out)
            loss.backward()
            acc = ref_accuracy(out.detach())
            # ---- oracle restatements vs the reference, same inputs
            o = oracle.infonce_step(q.detach(), k, my_mem, my_idx, T, all_k=all_k)
            my_idx = o["index"]
            assert my_idx == moco.index, (my_idx, moco.index)
            assert torch.equal(my_mem, moco.memory), "queue contents differ from the reference"
            assert torch.equal(labels, torch.zeros(B, dtype=torch.long))
            assert abs(float(o["loss"]) - float(loss)) <= 2e-6 * abs(float(loss)), (o["loss"], loss)
            assert torch.allclose(o["dq"], q.grad, rtol=1e-4, atol=1e-8)
            lg = oracle.logits_full(q.detach(), k, torch.from_numpy(rec["memory_before"]) if st == 0 else prev_mem, T)
            assert torch.allclose(lg, out.detach(), rtol=1e-6, atol=1e-6)
            myacc = [float(a) for a in oracle.topk_accuracy(out.detach())]
            assert myacc == acc, (myacc, acc)
            rk = oracle.positive_rank(out.detach())
            assert float((rk < 1).float().mean() * 100) == acc[0] and float((rk < 5).float().mean() * 100) == acc[1]
            prev_mem = moco.memory.clone()
            rec.update({
                f"q{st}": q.detach().numpy().copy(), f"k{st}": k.numpy().copy(),
                f"loss{st}": np.float64(float(loss)), f"dq{st}": q.grad.numpy().copy(),
                f"logits_head{st}": out.detach()[:, :9].numpy().copy(),
                f"lse{st}": torch.logsumexp(out.detach().double(), 1).numpy().copy(),
                f"rank{st}": rk.numpy().copy(), f"acc{st}": np.array(acc),
                f"index_after{st}": np.int64(moco.index),
                f"memory_sum_after{st}": np.float64(float(moco.memory.double().sum())),
            })
            if all_k is not None:
                rec[f"all_k{st}"] = all_k.numpy().copy()
        rec["memory_after"] = moco.memory.numpy().copy()
        rec["steps"] = np.int64(steps)
        return rec

    # small case with every input stored (K=256): 3 consecutive steps, pointer advances 0->24
    small = run_moco_case(seed=11, B=8, K=256, d=128, T=0.07, steps=3)
    np.savez_compressed(os.path.join(args.out, "infonce_small.npz"), **small)
    # wrap-around: pointer 236 + 40 gathered rows > 256  (SURVEY Appendix C "wrap test", scaled)
    wrap = run_moco_case(seed=12, B=8, K=256, d=128, T=0.07, n_all=40, start_index=236, steps=2)
    np.savez_compressed(os.path.join(args.out, "infonce_wrap.npz"), **wrap)
    # other feature dims
    d64 = run_moco_case(seed=13, B=5, K=192, d=64, T=0.2, steps=1)
    np.savez_compressed(os.path.join(args.out, "infonce_d64.npz"), **d64)

    # config 1: inputs by seed, outputs stored (queue 4096x128 is
    # 2 MB: not stored; tests regenerate it from the seed and verify its checksum)
    c1 = run_moco_case(seed=1, B=32, K=4096, d=128, T=0.07, steps=1)
    assert abs(c1["loss0"] - 8.9238758087) < 1e-5, c1["loss0"]
    keep = {k: v for k, v in c1.items() if k not in ("memory_before", "memory_after")}
    keep["memory_before_sum"] = np.float64(c1["memory_before"].astype(np.float64).sum())
    keep["memory_before_head"] = c1["memory_before"][:4].copy()
    np.savez_compressed(os.path.join(args.out, "infonce_c1.npz"), **keep)

    # ------------------------------------------------------------------ graph head cases
    def run_graph_case(seed_w, seed_u, shape, sub_sample, max_hop=3, alpha=0.5, temperature=1.0):
        B, C, T, H, W = shape
        torch.manual_seed(seed_w)
        mod = tg.TemporalGraphAug(C, sub_sample=sub_sample, max_hop=max_hop, alpah=alpha,
                                  temperature=temperature)
        x = torch.randn(*shape).requires_grad_(True)
        dy = torch.randn(*shape)
        torch.manual_seed(seed_u)
        y = mod(x)
        y.backward(dy)
        torch.manual_seed(seed_u)
        u = torch.rand(B, T, T)          # the first RNG draw inside forward (rsample)
        conv_q = mod.g_q[0] if sub_sample else mod.g_q
        conv_k = mod.g_k[0] if sub_sample else mod.g_k
        wq, wk, wg = conv_q.weight.detach(), conv_k.weight.detach(), mod.gcns[0].conv.weight.detach()
        # oracle vs reference: forward bit-equal, hop graph equal, backward to 1e-5
        my_y, sim, adj, s = oracle.graph_forward(x.detach(), wq, wk, wg, u, alpha, max_hop, temperature,
                                                 sub_sample=sub_sample)
        assert torch.equal(my_y, y.detach()), float((my_y - y.detach()).abs().max())
        hop_ref = tg.TemporalGraph(tem_len=T, max_hop=max_hop).temporal_graph
        hop_me = torch.from_numpy(oracle.hop_distance(T, max_hop)).float()
        hop_me[hop_me < 0] = float("inf")
        assert torch.equal(hop_ref, hop_me)
        _, dx, dwq, dwk, dwg = oracle.graph_forward_backward(x.detach(), wq, wk, wg, u, dy, alpha=alpha,
                                                             max_hop=max_hop, temperature=temperature,
                                                             sub_sample=sub_sample)
        for a, b in ((dx, x.grad), (dwq, conv_q.weight.grad), (dwk, conv_k.weight.grad),
                     (dwg, mod.gcns[0].conv.weight.grad)):
            assert torch.allclose(a, b, rtol=1e-4, atol=1e-6), float((a - b).abs().max())
        # closed-form core backward vs autograd of the core
        gq = og._project(x.detach(), wq, sub_sample, True).reshape(B, wq.shape[0], T, -1).requires_grad_(True)
        gk = og._project(x.detach(), wk, sub_sample, True).reshape(B, wk.shape[0], T, -1).requires_grad_(True)
        sup = F.conv3d(x.detach(), wg).reshape(B, C, T, -1).requires_grad_(True)
        yy, sim2, adj2, s2 = og.graph_core(gq, gk, sup, u, alpha, max_hop, temperature)
        g_auto = torch.autograd.grad(yy, [gq, gk, sup], dy.reshape(B, C, T, -1))
        g_mine = og.graph_core_backward(gq.detach(), gk.detach(), sup.detach(), sim2.detach(), adj2.detach(),
                                        s2.detach(), dy.reshape(B, C, T, -1), alpha, max_hop, temperature)
        for a, b in zip(g_mine, g_auto):
            assert torch.allclose(a, b, rtol=2e-4, atol=1e-6), float((a - b).abs().max())
        return {
            "x": x.detach().numpy(), "dy": dy.numpy(), "u": u.numpy(),
            "wq": wq.numpy(), "wk": wk.numpy(), "wg": wg.numpy(),
            "y": y.detach().numpy(), "sim": sim.numpy(), "adj": adj.numpy(), "s": s.numpy(),
            "dx": x.grad.numpy(), "dwq": conv_q.weight.grad.numpy(), "dwk": conv_k.weight.grad.numpy(),
            "dwg": mod.gcns[0].conv.weight.grad.numpy(),
            "d_gq": g_auto[0].numpy(), "d_gk": g_auto[1].numpy(), "d_support": g_auto[2].numpy(),
            "sub_sample": np.int64(sub_sample), "max_hop": np.int64(max_hop), "alpha": np.float64(alpha),
            "temperature": np.float64(temperature),
            "hop": oracle.hop_distance(T, max_hop), "theta": oracle.hop_weights(max_hop, alpha),
        }

    # config 1 graph: x=[8,128,4,1,1], sub_sample=False
    torch.manual_seed(3)
    g1 = run_graph_case(3, 7, (8, 128, 4, 1, 1), sub_sample=False)
    np.savez_compressed(os.path.join(args.out, "graph_c1.npz"), **g1)
    # feature-map case with spatial pooling, T=8 (> max_hop + 1 so the hop mask bites)
    g2 = run_graph_case(4, 8, (2, 16, 8, 4, 4), sub_sample=True)
    np.savez_compressed(os.path.join(args.out, "graph_fmap.npz"), **g2)
    # odd sizes: T=5, C=12 (C'=6), 3x3 spatial without pooling, max_hop=1, temperature 0.5
    g3 = run_graph_case(5, 9, (3, 12, 5, 3, 3), sub_sample=False, max_hop=1, alpha=0.3, temperature=0.5)
    np.savez_compressed(os.path.join(args.out, "graph_odd.npz"), **g3)
    # max_hop >= T (no masking at all), T=2
    g4 = run_graph_case(6, 10, (4, 8, 2, 1, 1), sub_sample=False, max_hop=3)
    np.savez_compressed(os.path.join(args.out, "graph_t2.npz"), **g4)

    # ------------------------------------------------------------------ SimSiam D
    torch.manual_seed(21)
    p = torch.randn(16, 64).requires_grad_(True)
    z = torch.randn(16, 64)
    loss = D()(p, z)
    loss.backward()
    assert abs(float(oracle.neg_cosine(p.detach(), z)) - float(loss)) < 1e-7
    assert torch.allclose(oracle.neg_cosine_grad(p.detach(), z), p.grad, rtol=1e-4, atol=1e-8)
    np.savez_compressed(os.path.join(args.out, "negcos.npz"), p=p.detach().numpy(), z=z.numpy(),
                        loss=np.float64(float(loss)), dp=p.grad.numpy())

    # ------------------------------------------------------------------ retrieval (video_retrieval.py:183-197)
    from sklearn.metrics.pairwise import cosine_distances

    def ref_retrieval(val, train, val_cls, train_cls):
        distances = cosine_distances(val, train)
        indices = np.argsort(distances)
        hits = {}
        for k in oracle.retrieval.KS:
            c = 0
            for ind, lab in zip(indices[:, :k], val_cls):
                if lab in train_cls[ind]:
                    c += 1
            hits[k] = c
        return hits, indices

    rng = np.random.default_rng(0)
    gal = rng.standard_normal((13320, 512)).astype(np.float32)
    qry = rng.standard_normal((3783, 512)).astype(np.float32)
    gl = rng.integers(0, 101, 13320)
    ql = rng.integers(0, 101, 3783)
    hits, _ = ref_retrieval(qry, gal, ql, gl)
    assert hits == {1: 28, 5: 151, 10: 349, 20: 666, 50: 1478}, hits
    idx, _ = oracle.cosine_topk(qry, gal, 50)
    assert oracle.recall_hits(idx, ql, gl) == hits
    rng = np.random.default_rng(5)
    sg = rng.standard_normal((300, 32)).astype(np.float32)
    sq = rng.standard_normal((40, 32)).astype(np.float32)
    sgl, sql = rng.integers(0, 7, 300), rng.integers(0, 7, 40)
    shits, sind = ref_retrieval(sq, sg, sql, sgl)
    sidx, _ = oracle.cosine_topk(sq, sg, 50)
    assert oracle.recall_hits(sidx, sql, sgl) == shits
    np.savez_compressed(os.path.join(args.out, "retrieval.npz"),
                        c5_hits=np.array([hits[k] for k in oracle.retrieval.KS]),
                        small_gallery=sg, small_queries=sq, small_gallery_labels=sgl, small_query_labels=sql,
                        small_hits=np.array([shits[k] for k in oracle.retrieval.KS]),
                        small_top10=sind[:, :10])
    print("golden fixtures written to", args.out)
    for f in sorted(os.listdir(args.out)):
        print("  %-24s %8d bytes" % (f, os.path.getsize(os.path.join(args.out, f))))


if __name__ == "__main__":
    main()
