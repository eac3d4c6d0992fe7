"""CPU: the oracle restatements against the fixtures generated from the reference (oracle/gen_golden.py)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle
from oracle import graph as og


def T_(a):
    return torch.from_numpy(np.asarray(a))


@pytest.mark.parametrize("name", ["infonce_small", "infonce_wrap", "infonce_d64"])
def test_infonce_steps_match_reference(golden, name):
    g = golden(name)
    mem = T_(g["memory_before"]).clone()
    idx = int(g["start_index"])
    T = float(g["T"])
    for st in range(int(g["steps"])):
        q, k = T_(g[f"q{st}"]), T_(g[f"k{st}"])
        all_k = T_(g[f"all_k{st}"]) if f"all_k{st}" in g else None
        lg = oracle.logits_full(q, k, mem, T)
        np.testing.assert_allclose(lg[:, :9].numpy(), g[f"logits_head{st}"], rtol=1e-6, atol=1e-6)
        o = oracle.infonce_step(q, k, mem, idx, T, all_k=all_k)
        idx = o["index"]
        assert idx == int(g[f"index_after{st}"])                                   # integer: exact
        assert abs(float(o["loss"]) - float(g[f"loss{st}"])) <= 1e-5 * abs(float(g[f"loss{st}"]))
        np.testing.assert_allclose(o["dq"].numpy(), g[f"dq{st}"], rtol=1e-4, atol=1e-8)
        np.testing.assert_array_equal(o["rank"].numpy(), g[f"rank{st}"])
        np.testing.assert_allclose(o["lse"].double().numpy(), g[f"lse{st}"], rtol=1e-6)
        assert float(mem.double().sum()) == pytest.approx(float(g[f"memory_sum_after{st}"]), abs=1e-9)
    assert torch.equal(mem, T_(g["memory_after"]))                                  # slot contents: exact


def test_wrap_around_slots(golden):
    g = golden("infonce_wrap")
    assert int(g["start_index"]) == 236 and int(g["index_after0"]) == (236 + 40) % 256
    mem0, mem1 = g["memory_before"], g["memory_after"]
    all_k0 = g["all_k0"]
    # after step 0 rows 236..255 hold all_k0[:20], rows 0..19 hold all_k0[20:] (then step 1 overwrites 20..59)
    np.testing.assert_array_equal(mem1[236:], all_k0[:20])
    np.testing.assert_array_equal(mem1[:20], all_k0[20:])
    np.testing.assert_array_equal(mem1[60:236], mem0[60:236])
    np.testing.assert_array_equal(oracle.ring_slots(236, 40, 256)[[0, 19, 20, 39]], [236, 255, 0, 19])


def test_config1_by_seed(golden):
    """inputs regenerated from the seed, outputs from the reference."""
    g = golden("infonce_c1")
    torch.manual_seed(1)
    mem = F.normalize(torch.randn(4096, 128))
    q = F.normalize(torch.randn(32, 128))
    k = F.normalize(torch.randn(32, 128))
    assert float(mem.double().sum()) == pytest.approx(float(g["memory_before_sum"]), abs=1e-9), "RNG stream changed"
    np.testing.assert_array_equal(mem[:4].numpy(), g["memory_before_head"])
    np.testing.assert_array_equal(q.numpy(), g["q0"])
    o = oracle.infonce_step(q, k, mem, 0, 0.07)
    assert float(o["loss"]) == pytest.approx(8.9238758087, rel=1e-6)
    assert float(o["dq"].abs().sum()) == pytest.approx(130.3457336426, rel=1e-5)
    np.testing.assert_allclose(o["dq"][0, :3].numpy(), [-0.0333443545, -0.0277747400, -0.0518033244], rtol=1e-5)
    assert o["index"] == 32 and torch.equal(mem[:32], k)
    assert float(mem.double().sum()) == pytest.approx(-3.7303031141, abs=1e-6)
    np.testing.assert_array_equal(g["acc0"], [0.0, 0.0])


def test_reference_head_step_port_matches_closed_form(golden):
    g = golden("infonce_small")
    mem_a, mem_b = T_(g["memory_before"]).clone(), T_(g["memory_before"]).clone()
    q = T_(g["q0"]).clone().requires_grad_(True)
    k = T_(g["k0"])
    loss, dq, idx, acc = oracle.infonce.reference_head_step(q, k, mem_a, 0, float(g["T"]))
    o = oracle.infonce_step(q.detach(), k, mem_b, 0, float(g["T"]))
    assert float(loss) == pytest.approx(float(o["loss"]), rel=1e-6)
    np.testing.assert_allclose(dq.numpy(), o["dq"].numpy(), rtol=1e-4, atol=1e-8)
    assert idx == o["index"] and torch.equal(mem_a, mem_b)
    assert [float(a) for a in acc] == list(g["acc0"])


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_sharded_merge_equals_unsharded(golden, world):
    g = golden("infonce_small")
    q, k, mem = T_(g["q0"]), T_(g["k0"]), T_(g["memory_before"])
    full = oracle.infonce_step(q, k, mem.clone(), 0, float(g["T"]))
    sh = oracle.sharded_infonce(q, k, mem, float(g["T"]), world)
    assert float(sh["loss"]) == pytest.approx(float(full["loss"]), rel=2e-6)
    np.testing.assert_array_equal(sh["rank"].numpy(), full["rank"].numpy())
    np.testing.assert_allclose(sh["dq"].numpy(), full["dq"].numpy(), rtol=2e-4, atol=1e-8)


def test_sharded_enqueue_ownership():
    torch.manual_seed(0)
    K, d, W = 64, 8, 4
    full = torch.randn(K, d)
    shards = [full[r * 16:(r + 1) * 16].clone() for r in range(W)]
    keys = torch.randn(24, d)
    idx_full = oracle.enqueue(full, keys, 54)
    for r in range(W):
        assert oracle.ring.enqueue_sharded(shards[r], keys, 54, K, r * 16) == idx_full
    assert torch.equal(torch.cat(shards), full) and idx_full == (54 + 24) % 64


def test_enqueue_rejects_more_rows_than_slots():
    with pytest.raises(ValueError):
        oracle.enqueue(torch.zeros(4, 8), torch.zeros(5, 8), 0)


@pytest.mark.parametrize("name", ["graph_c1", "graph_fmap", "graph_odd", "graph_t2"])
def test_graph_forward_backward_match_reference(golden, name):
    g = golden(name)
    kw = dict(alpha=float(g["alpha"]), max_hop=int(g["max_hop"]), temperature=float(g["temperature"]),
              sub_sample=bool(g["sub_sample"]))
    x, wq, wk, wg, u, dy = (T_(g[n]) for n in ("x", "wq", "wk", "wg", "u", "dy"))
    y, sim, adj, s = oracle.graph_forward(x, wq, wk, wg, u, **kw)
    assert torch.equal(y, T_(g["y"]))                                               # bit-equal to the reference
    np.testing.assert_array_equal(s.numpy(), g["s"])
    np.testing.assert_array_equal(oracle.hop_distance(x.shape[2], kw["max_hop"]), g["hop"])
    _, dx, dwq, dwk, dwg = oracle.graph_forward_backward(x, wq, wk, wg, u, dy, **kw)
    for a, n in ((dx, "dx"), (dwq, "dwq"), (dwk, "dwk"), (dwg, "dwg")):
        np.testing.assert_allclose(a.numpy(), g[n], rtol=1e-4, atol=1e-6)
    # closed-form core backward (what the CUDA kernel implements) vs the reference's autograd
    B, C, T = x.shape[:3]
    gq = og._project(x, wq, kw["sub_sample"], True).reshape(B, wq.shape[0], T, -1)
    gk = og._project(x, wk, kw["sub_sample"], True).reshape(B, wk.shape[0], T, -1)
    sup = F.conv3d(x, wg).reshape(B, C, T, -1)
    d_gq, d_gk, d_sup = og.graph_core_backward(gq, gk, sup, sim, adj, s, dy.reshape(B, C, T, -1), kw["alpha"],
                                               kw["max_hop"], kw["temperature"])
    np.testing.assert_allclose(d_gq.numpy(), g["d_gq"], rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(d_gk.numpy(), g["d_gk"], rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(d_sup.numpy(), g["d_support"], rtol=2e-4, atol=1e-6)


def test_graph_config1_known_answers(golden):
    """SURVEY Appendix C G2."""
    g = golden("graph_c1")
    assert float(g["y"].astype(np.float64).sum()) == pytest.approx(54.6891079033, abs=1e-4)
    np.testing.assert_allclose(g["y"][0, 0, :, 0, 0], [-0.4872, 0.2106, -0.0156, -0.5280], atol=5e-5)
    np.testing.assert_allclose(g["theta"], [1.0, 0.8240271, 0.6329011, 0.5496640], rtol=1e-6)


def test_negcos_matches_reference(golden):
    g = golden("negcos")
    p, z = T_(g["p"]), T_(g["z"])
    assert float(oracle.neg_cosine(p, z)) == pytest.approx(float(g["loss"]), abs=1e-7)
    np.testing.assert_allclose(oracle.neg_cosine_grad(p, z).numpy(), g["dp"], rtol=1e-4, atol=1e-8)


def test_retrieval_small_matches_reference(golden):
    g = golden("retrieval")
    idx, _ = oracle.cosine_topk(g["small_queries"], g["small_gallery"], 50)
    np.testing.assert_array_equal(idx[:, :10], g["small_top10"])
    hits = oracle.recall_hits(idx, g["small_query_labels"], g["small_gallery_labels"])
    assert [hits[k] for k in oracle.retrieval.KS] == list(g["small_hits"])
    assert list(g["c5_hits"]) == [28, 151, 349, 666, 1478]


def test_projection_tail_oracle_matches_reference_fixture(golden):
    """oracle.infonce.head_from_projections against what the reference's own Normalize + RGBMoCo + NCESoftmaxLoss + autograd
    produced (tests/golden/proj_tail.npz, written by oracle/gen_golden_proj.py)."""
    import torch
    from oracle.infonce import head_from_projections
    g = golden("proj_tail")
    zq, zk, mem = torch.from_numpy(g["zq"]), torch.from_numpy(g["zk"]), torch.from_numpy(g["memory_before"]).clone()
    o = head_from_projections(zq, zk, mem, 0, float(g["T"]))
    assert abs(float(o["loss"]) - float(g["loss"])) <= 2e-6 * float(g["loss"])
    assert float((o["dz"] - torch.from_numpy(g["dz"])).abs().max()) <= 1e-4 * float(np.abs(g["dz"]).max())
    assert o["index"] == int(g["index_after"])
    assert torch.equal(mem[:zq.shape[0]], torch.from_numpy(g["enqueued_rows"]))        # slot contents: exact


def test_cmc_moco_and_jig_heads_match_reference(golden):
    """Fixtures from the reference's CMCMoCo / RGBMoCo(q_jig=...) (oracle/gen_golden_cmc.py): every head is the ordinary
    InfoNCE step against the OTHER modality's keys and queue (mem_moco.py:120-125); both queues take the same slots."""
    g = golden("cmc_moco")
    K, T = int(g["K"]), float(g["T"])
    m1, m2, idx = T_(g["memory_1_before"]).clone(), T_(g["memory_2_before"]).clone(), int(g["start_index"])
    for st in range(int(g["steps"])):
        k1, k2 = T_(g[f"k1_{st}"]), T_(g[f"k2_{st}"])
        nheads = 4 if f"q{st}_3" in g else 2
        for i in range(nheads):
            k, mem = ((k2, m2), (k1, m1))[i % 2]
            o = oracle.infonce_step(T_(g[f"q{st}_{i}"]), k, mem.clone(), 0, T)
            assert abs(float(o["loss"]) - float(g[f"loss{st}_{i}"])) <= 1e-5 * abs(float(g[f"loss{st}_{i}"]))
            np.testing.assert_allclose(o["dq"].numpy(), g[f"dq{st}_{i}"], rtol=1e-4, atol=1e-7)
        new1 = T_(g[f"all_k1_{st}"]) if f"all_k1_{st}" in g else k1
        new2 = T_(g[f"all_k2_{st}"]) if f"all_k2_{st}" in g else k2
        oracle.enqueue(m1, new1, idx)
        idx = oracle.enqueue(m2, new2, idx)
        assert idx == int(g[f"index_after{st}"])
    assert idx < 20                                                       # the ring wrapped
    assert torch.equal(m1, T_(g["memory_1_after"])) and torch.equal(m2, T_(g["memory_2_after"]))
    j = golden("rgb_jig")
    mem, idx = T_(j["memory_before"]).clone(), 0
    for st in range(int(j["steps"])):
        k = T_(j[f"k{st}"])
        for name in ("", "_jig"):
            o = oracle.infonce_step(T_(j[f"q{name}{st}"]), k, mem.clone(), 0, float(j["T"]))
            assert abs(float(o["loss"]) - float(j[f"loss{name}{st}"])) <= 1e-5 * abs(float(j[f"loss{name}{st}"]))
            np.testing.assert_allclose(o["dq"].numpy(), j[f"dq{name}{st}"], rtol=1e-4, atol=1e-7)
        idx = oracle.enqueue(mem, k, idx)
        assert idx == int(j[f"index_after{st}"])
    assert torch.equal(mem, T_(j["memory_after"]))


@pytest.mark.parametrize("name", ["proj", "pred"])
def test_simsiam_mlp_restatement_matches_reference(golden, name):
    """oracle.mlp (Linear + BatchNorm1d [+ ReLU] forward / backward by hand, fp64) against the outputs, gradients and running
    statistics of the reference's ProjectionMLP / PredictionMLP (oracle/gen_golden_mlp.py)."""
    from oracle import mlp as om
    g = golden("simsiam_mlp")
    p = {k[len(name) + 8:]: np.asarray(v, dtype=np.float64) for k, v in g.items() if k.startswith(name + ".before.")}
    x, w = np.asarray(g[name + ".x"], np.float64), np.asarray(g[name + ".w"], np.float64)
    if name == "proj":
        y, caches = om.projection_mlp(x, p)
        dx, grads = om.projection_mlp_backward(w, caches)
    else:
        y, caches = om.prediction_mlp(x, p)
        dx, grads = om.prediction_mlp_backward(w, caches)
    np.testing.assert_allclose(y, g[name + ".y"], rtol=0, atol=2e-5 * np.abs(y).max())
    np.testing.assert_allclose(dx, g[name + ".dx"], rtol=0, atol=1e-4 * np.abs(dx).max())
    for i, blk in enumerate(("l1", "l2", "l3")[:len(grads) if name == "proj" else 1]):
        gr = grads[i]
        for key, ref in (("dW", blk + ".0.weight"), ("db", blk + ".0.bias"), ("dgamma", blk + ".1.weight"), ("dbeta", blk + ".1.bias")):
            r = g[name + ".grad." + ref]
            # (the bias in front of a BatchNorm has a mathematically zero gradient: fp32 autograd leaves ~1e-5 of noise there)
            np.testing.assert_allclose(gr[key], r, rtol=0, atol=1e-4 if key == "db" else 1e-4 * max(np.abs(r).max(), 1e-3))
        c = caches[i]
        np.testing.assert_allclose(0.9 * p[blk + ".1.running_mean"] + 0.1 * c["mean"], g[name + ".after." + blk + ".1.running_mean"], atol=1e-5)
        np.testing.assert_allclose(0.9 * p[blk + ".1.running_var"] + 0.1 * c["var_unb"], g[name + ".after." + blk + ".1.running_var"], atol=1e-5)
        assert int(g[name + ".after." + blk + ".1.num_batches_tracked"]) == 1


def test_instance_bank_restatement_matches_reference(golden):
    """oracle/bank.py against tests/golden/bank.npz (RGBMem / CMCMem / AliasMethod / NCECriterion of the reference run by
    oracle/gen_golden_bank.py): alias tables and the draw arithmetic exact, logits exact, gradients 1e-6, bank update exact
    (incl. the duplicated index: the last occurrence wins)."""
    from oracle import bank as ob
    g = golden("bank")
    prob, alias = ob.alias_tables(T_(g["alias_p"]))
    assert torch.equal(prob, T_(g["alias_prob"])) and torch.equal(alias, T_(g["alias_alias"]))
    assert torch.equal(ob.alias_pick(prob, alias, T_(g["alias_kk"]), T_(g["alias_b"])), T_(g["alias_draw"]))
    T, m = float(g["T"]), float(g["m"])
    mem = T_(g["rgb_memory_before"]).clone()
    for st in range(2):
        x, y, idx = T_(g["rgb%d_x" % st]), T_(g["rgb%d_y" % st]), T_(g["rgb%d_idx" % st])
        assert torch.equal(idx[:, 0], y)
        assert torch.equal(ob.bank_logits(x, mem, idx, T), T_(g["rgb%d_logits" % st]))
        dx = ob.bank_grad_x(T_(g["rgb%d_glogits" % st]), mem, idx, T)
        np.testing.assert_allclose(dx.numpy(), g["rgb%d_dx" % st], rtol=1e-5, atol=1e-7)
        if st == 1:
            assert torch.equal(ob.bank_logits(T_(g["rgb1_x_jig"]), mem, idx, T), T_(g["rgb1_logits_jig"]))
            ob.bank_update(mem, T_(g["rgb1_all_x"]), T_(g["rgb1_all_y"]), m)
        else:
            ob.bank_update(mem, x, y, m)
        assert torch.equal(mem, T_(g["rgb%d_memory_after" % st]))
    np.testing.assert_allclose(float(ob.nce_criterion(T_(g["nce_x"]), int(g["n_data"]))), float(g["nce_loss"]), rtol=1e-6)
