"""CPU: host-side mirror of the reference interfaces (no kernels run here)."""
import types

import numpy as np
import pytest
import torch
import torch.nn as nn

import gca_b200
from gca_b200.memory.moco_queue import FusedLogits


def cfg(mem_type="moco", crit="crossentropy", K=256, **extra):
    c = types.SimpleNamespace(
        CONTRAST=types.SimpleNamespace(MEM_TYPE=mem_type, NCE_K=K, NCE_T=0.07, NCE_M=0.5, **extra),
        CROSS=types.SimpleNamespace(MODALITY="visual", FEAT_DIM=128, CRITERION=crit))
    return c


def test_factories_follow_reference_contract():
    m = gca_b200.create_contrast(cfg(), n_data=1000)
    assert isinstance(m, gca_b200.RGBMoCo) and m.K == 256 and m.T == 0.07 and m.index == 0
    assert list(m.state_dict().keys()) == ["memory"] and m.memory.shape == (256, 128) and m.memory.dtype == torch.float32
    assert gca_b200.create_contrast(cfg("simsiam"), 1) is None                       # lib/memory/build.py:13-14
    assert isinstance(gca_b200.create_criterion(cfg(), 1), gca_b200.NCESoftmaxLoss)
    assert isinstance(gca_b200.create_criterion(cfg(crit="simsiam_d"), 1), gca_b200.D)
    with pytest.raises(NotImplementedError):
        gca_b200.create_contrast(cfg("nope"), 1)
    with pytest.raises(NotImplementedError):
        gca_b200.create_criterion(cfg(crit="nope"), 1)
    b = gca_b200.create_contrast(cfg(QUEUE_DTYPE="bf16"), 1)
    assert b.memory.dtype == torch.bfloat16
    # the instance-bank branch (lib/memory/build.py:6-9, 24-25)
    bank = gca_b200.create_contrast(cfg("bank", K=32), n_data=50)
    assert isinstance(bank, gca_b200.RGBMem) and (bank.K, bank.T, bank.m) == (32, 0.07, 0.5)
    assert list(bank.state_dict().keys()) == ["memory"] and bank.memory.shape == (50, 128)
    c2 = cfg("bank", K=32)
    c2.CROSS.MODALITY = "cross"
    two = gca_b200.create_contrast(c2, n_data=50)
    assert isinstance(two, gca_b200.CMCMem) and list(two.state_dict().keys()) == ["memory_1", "memory_2"]
    assert isinstance(gca_b200.create_criterion(cfg(crit="NCE"), 77), gca_b200.NCECriterion)
    with pytest.raises(RuntimeError, match="no CPU path"):
        bank(torch.randn(4, 128), torch.arange(4))


def test_alias_sampler_and_nce_criterion_match_reference_fixture(golden):
    """AliasMethod tables / draw arithmetic and NCECriterion of the drop-in against tests/golden/bank.npz (integer tables and
    samples exact)."""
    import numpy as np
    from gca_b200.memory import AliasMethod
    g = golden("bank")
    am = AliasMethod(torch.from_numpy(g["alias_p"]))
    assert torch.equal(am.prob, torch.from_numpy(g["alias_prob"])) and torch.equal(am.alias, torch.from_numpy(g["alias_alias"]))
    torch.manual_seed(4)                                                   # the seed the fixture's reference draw used
    assert torch.equal(am.draw(1000), torch.from_numpy(g["alias_draw"]))
    uni = AliasMethod(torch.ones(9))
    assert torch.equal(uni.prob, torch.ones(9))
    loss = gca_b200.NCECriterion(int(g["n_data"]))(torch.from_numpy(g["nce_x"]))
    np.testing.assert_allclose(float(loss), float(g["nce_loss"]), rtol=1e-6)


def test_queue_init_consumes_rng_like_reference():
    torch.manual_seed(1)
    m = gca_b200.RGBMoCo(128, K=4096)
    torch.manual_seed(1)
    ref = torch.nn.functional.normalize(torch.randn(4096, 128))                       # mem_moco.py:57-58
    assert torch.equal(m.memory, ref)
    assert torch.allclose(m.memory.norm(dim=1), torch.ones(4096), atol=1e-6)


def test_bf16_queue_checkpoints_as_fp32():
    torch.manual_seed(0)
    m = gca_b200.RGBMoCo(32, K=64, queue_dtype="bf16")
    sd = m.state_dict()
    assert sd["memory"].dtype == torch.float32 and torch.equal(sd["memory"].to(torch.bfloat16), m.memory)
    m2 = gca_b200.RGBMoCo(32, K=64, queue_dtype="bf16")
    m2.load_state_dict(sd)
    assert m2.memory.dtype == torch.bfloat16 and torch.equal(m2.memory, m.memory)
    ref_style = {"memory": torch.nn.functional.normalize(torch.randn(64, 32))}        # an upstream fp32 checkpoint
    m2.load_state_dict(ref_style)
    assert torch.equal(m2.memory, ref_style["memory"].to(torch.bfloat16))


def test_pointer_bookkeeping():
    m = gca_b200.RGBMoCo(32, K=64)
    m.index = 60
    m._update_pointer(10)
    assert m.index == 6                                                               # mem_moco.py:14-15


def test_fused_logits_topk_drives_reference_accuracy():
    """`accuracy` (metric.py:44-67, reshape-fixed) on the handle must equal accuracy on real logits."""
    import oracle
    torch.manual_seed(3)
    lg = torch.randn(64, 257)
    lg[:8, 0] += 5.0
    rank = oracle.positive_rank(lg)
    h = FusedLogits(torch.tensor(0.), torch.zeros(64), torch.zeros(64), lg[:, 0], rank.int(), 64, 256)
    assert h.shape[0] == 64 and h.size(1) == 257 and h.detach().shape == h.shape
    _, pred = h.topk(5, 1, True, True)
    correct = pred.t().eq(torch.zeros(1, 64, dtype=pred.dtype))
    mine = [float(correct[:k].reshape(-1).float().sum() * (100.0 / 64)) for k in (1, 5)]
    ref = [float(a) for a in oracle.topk_accuracy(lg, (1, 5))]
    assert mine == ref and ref[0] > 0
    crit = gca_b200.NCESoftmaxLoss()
    assert crit(h) is h.loss
    assert float(crit(lg)) == pytest.approx(float(oracle.infonce_loss(lg)), rel=1e-6)


def test_graph_module_matches_reference_layout_and_init(golden):
    g = golden("graph_c1")
    torch.manual_seed(3)
    m = gca_b200.TemporalGraphAug(128, sub_sample=False)
    # same parameter names, shapes and -- for the same seed -- the same values as the (shimmed) reference ctor
    sd = m.state_dict()
    assert sorted(sd) == ["g_k.weight", "g_q.weight", "gcns.0.conv.weight"]
    np.testing.assert_array_equal(sd["g_q.weight"].numpy(), g["wq"])
    np.testing.assert_array_equal(sd["g_k.weight"].numpy(), g["wk"])
    np.testing.assert_array_equal(sd["gcns.0.conv.weight"].numpy(), g["wg"])
    m2 = gca_b200.TemporalGraphAug(16)                         # sub_sample=True default -> Sequential(conv, pool)
    assert sorted(m2.state_dict()) == ["g_k.0.weight", "g_q.0.weight", "gcns.0.conv.weight"]
    assert m2.inter_channels == 8 and m2.gcns[0].conv.out_channels == 16 and m2.alpha == 0.5 and m2.max_hop == 3
    # mask_frame=True: upstream constructs, but no forward can complete (probed on the reference: its mask loop indexes the
    # batch axis, temporal_graph.py:169-174) -- IndexError when B < nei_size (default T), else NaN rows -> ValueError from
    # RelaxedBernoulli's argument check.  The drop-in shows the same error behaviour.
    m3 = gca_b200.TemporalGraphAug(16, sub_sample=False, mask_frame=True)
    with pytest.raises(IndexError):
        m3(torch.randn(3, 16, 8, 2, 2))
    with pytest.raises(ValueError):
        m3(torch.randn(8, 16, 8, 2, 2))
    with pytest.raises(ValueError):
        gca_b200.TemporalGraphAug(16, sub_sample=False, mask_frame=True, nei_size=4)(torch.randn(4, 16, 8, 2, 2))
    with pytest.raises(NotImplementedError):
        gca_b200.TemporalGraphAug(16, num_gcn_layers=2)


def test_build_aug_block_wraps_named_modules():
    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.base = nn.Sequential(nn.Conv3d(3, 8, 1), nn.ReLU(), nn.Conv3d(8, 16, 1),
                                      nn.Sequential(nn.ReLU(), nn.Conv3d(16, 4, 1)))
            self.layer2 = nn.Conv3d(4, 4, 1)

    net = gca_b200.build_aug_block(Net(), ["base.2", "base.3", "layer2"], n_segments=16)
    for mod, cin in ((net.base[2], 8), (net.base[3], 16), (net.layer2, 4)):
        assert isinstance(mod, nn.Sequential) and isinstance(mod[0], gca_b200.TemporalGraphAug)
        assert mod[0].in_channels == cin
    assert isinstance(net.base[0], nn.Conv3d)
    agg = gca_b200.get_agg("avg", "3D")
    assert agg(torch.ones(2, 3, 4)).shape == (2, 4)            # upstream always pools dim 1 (build.py:6)


# --------------------------------------------------------------------------------------------- ShuffleBN routing tables
@pytest.mark.parametrize("world,bsz", [(1, 5), (2, 6), (4, 3), (8, 16)])
def test_shuffle_plan_reproduces_reference_indexing(world, bsz):
    """gca_b200.dist.shuffle_plan against the reference's gather-everything indexing (train...:203-215), all ranks
    simulated in one process: what every rank sends, concatenated by destination, is exactly node_x[this_ids]."""
    from gca_b200.dist import shuffle_plan
    from oracle.shuffle import shuffle_bn_all_ranks
    gen = torch.Generator().manual_seed(world * 100 + bsz)
    xs = [torch.randn(bsz, 7, generator=gen) for _ in range(world)]
    ids = torch.randperm(world * bsz, generator=gen)
    _, _, ref_this = shuffle_bn_all_ranks(xs, lambda t: t, ids)
    plans = [shuffle_plan(ids, bsz, world, r) for r in range(world)]
    for r in range(world):
        rows, sc, rc, place = plans[r]
        assert sum(sc) == bsz and sum(rc) == bsz
        assert rc == [plans[s][1][r] for s in range(world)]               # what r receives from s == what s sends to r
        # emulate the all-to-all: the chunk rank s sends to r, in source-rank order
        recv = []
        for s in range(world):
            srows, ssc = plans[s][0], plans[s][1]
            off = sum(ssc[:r])
            recv.append(xs[s][srows[off:off + ssc[r]]])
        recv = torch.cat(recv)
        out = torch.empty_like(recv)
        out[place] = recv
        assert torch.equal(out, ref_this[r])


def test_ema_layout_rules_on_cpu():
    """MomentumUpdater refuses CPU parameters (no CPU path) and its density check accepts permuted-dense layouts only."""
    from gca_b200.ema import MomentumUpdater, _dense
    a = torch.randn(4, 3, 5, 6, 7)
    assert _dense(a) and _dense(a.to(memory_format=torch.channels_last_3d)) and _dense(a.permute(4, 3, 2, 1, 0))
    assert not _dense(a[:, :2]) and not _dense(torch.randn(6)[::2])
    assert _dense(torch.randn(1, 5)) and _dense(torch.randn(()))
    with pytest.raises(RuntimeError):
        MomentumUpdater(nn.Linear(3, 3), nn.Linear(3, 3))


def test_peer_exchange_and_graphed_steps_need_cuda():
    """The multi-GPU helpers have no CPU path either: they raise before touching the library."""
    from gca_b200.graphed import GraphedMoCoStep
    m = gca_b200.RGBMoCo(32, K=64)
    with pytest.raises(RuntimeError):
        GraphedMoCoStep(m, 8)


def test_shuffle_plan_properties_randomised():
    """Random worlds / batch sizes / permutations: counts are consistent across ranks, every row is sent exactly once,
    and `place` is a permutation (hypothesis-style sweep with a fixed seed so the suite stays deterministic)."""
    from gca_b200.dist import shuffle_plan
    rng = np.random.default_rng(123)
    for _ in range(60):
        world, bsz = int(rng.integers(1, 9)), int(rng.integers(1, 33))
        ids = torch.from_numpy(rng.permutation(world * bsz))
        plans = [shuffle_plan(ids, bsz, world, r) for r in range(world)]
        sent = []
        for r, (rows, sc, rc, place) in enumerate(plans):
            assert len(sc) == world and len(rc) == world and sum(sc) == sum(rc) == bsz
            assert sorted(rows.tolist()) == list(range(bsz))                    # every local row leaves exactly once
            assert sorted(place.tolist()) == list(range(bsz))
            for dst in range(world):
                assert sc[dst] == plans[dst][2][r]
            sent.append(rows)
        # reverse ids undo the shuffle: argsort(ids)[ids[g]] == g
        rev = torch.argsort(ids)
        assert torch.equal(ids[rev], torch.arange(world * bsz))


def test_pretrain_encoder_has_the_reference_r3d18_layer_shapes():
    """The encoder around the head in the pre-train clips/s figure (tools/pretrain_step.py) has exactly the parameter shapes and
    the feature-map shape of the reference's resnet18(sample_size=112, sample_duration=16) (backbone_3d/resnet.py:108-222);
    the list was taken from the reference's own class (oracle/gen_golden_r3d.py)."""
    import json
    import os
    import sys
    from conftest import ROOT
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import pretrain_step as ps
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "r3d18_shapes.json")))
    enc = ps.R3D18Shape()
    shapes = sorted(list(p.shape) for n, p in enc.named_parameters() if not n.startswith("head."))
    assert shapes == g["sorted_param_shapes"]
    assert sum(p.numel() for n, p in enc.named_parameters() if not n.startswith("head.")) == g["n_params"] == 33203904
    with torch.no_grad():
        f = enc.layers(enc.stem(torch.randn(1, 3, 16, 112, 112)))
    assert list(f.shape) == g["feature_map_shape_for_1x3x16x112x112"]


def test_cpu_placement_helper_is_best_effort():
    """gca_b200.affinity: without a CUDA device (or NVML / sysfs information) nothing is bound and nothing raises; restore() of
    an unbound record is a no-op; the sysfs parser understands cpulist ranges."""
    import os
    from gca_b200 import affinity
    before = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    rec = affinity.bind_cpu_to_device(0)
    assert isinstance(rec, dict) and set(rec) >= {"bound", "source", "cpus", "previous"}
    if not torch.cuda.is_available():
        assert rec["bound"] is False
    affinity.restore(rec)
    affinity.restore(None)
    if before is not None:
        assert os.sched_getaffinity(0) == before
    cpus, src = affinity._local_cpus("0000ffff:ff:1f.0")                # no such device: unknown, empty
    assert cpus == set() and src == "unknown"
