"""CPU, world_size 2 (gloo): the host-side orchestration of the K-sharded queue (gca_b200.dist.ShardedRGBMoCo):
gather of q / k, all-gather of the per-shard partials, combine, reduce-scatter, finish, sharded enqueue ownership and the
replicated ring pointer.  The four compute steps are injected from the oracle here (the product binds them to the CUDA
library); the same module runs over NCCL on the GPU box (bench.py --gpus N)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, PKG, ROOT


class OracleCompute(object):
    """Drop-in for gca_b200.dist.ShardCompute computing on CPU tensors with the oracle (TEST ONLY)."""

    def shard_fwd(self, q_all, k_all, shard, T, algo, want_grad):
        import oracle
        pos = (q_all * k_all).sum(1) / T
        m, s, cnt = oracle.lse_partials(q_all, shard, T, pos)
        stats = torch.stack([m, s, cnt.to(torch.int32).view(torch.float32)])          # count travels as raw bits
        lg = (q_all @ shard.float().t()) / T
        acc = torch.exp(lg - m[:, None]) @ shard.float() if want_grad else None
        return pos, stats, acc

    def shard_combine(self, all_stats, rank_id, pos, acc):
        import oracle
        ms, ss = all_stats[:, 0], all_stats[:, 1]
        cnt = all_stats[:, 2].contiguous().view(torch.int32).sum(0).to(torch.int32)
        lse = oracle.merge_partials(ms, ss, pos)
        if acc is not None:
            acc.mul_(torch.exp(ms[rank_id] - lse)[:, None])
        return lse, lse - pos, cnt

    def shard_finish(self, acc_loc, k_loc, pos_loc, lse_loc, loss_rows_loc, T):
        B = pos_loc.shape[0]
        dq = None
        if acc_loc is not None:
            dq = ((torch.exp(pos_loc - lse_loc) - 1)[:, None] * k_loc + acc_loc) / (T * B)
        return dq, loss_rows_loc.mean()

    def enqueue(self, shard, keys, index, K, k_begin):
        from oracle.ring import enqueue_sharded
        return enqueue_sharded(shard, keys, index, K, k_begin)


def _worker(rank, world, port, out_dir):
    for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from gca_b200.dist import ShardedRGBMoCo
    from gca_b200.memory.losses import NCESoftmaxLoss
    g = dict(np.load(os.path.join(GOLDEN, "infonce_small.npz")))
    K, d, T = 256, 128, float(g["T"])
    torch.manual_seed(100 + rank)                     # ranks draw DIFFERENT queues; rank 0's must win (train...:233-242)
    moco = ShardedRGBMoCo(d, K=K, T=T, compute=OracleCompute())
    torch.manual_seed(100)
    full0 = torch.nn.functional.normalize(torch.randn(K, d))
    Ks = K // world
    assert torch.equal(moco.memory, full0[rank * Ks:(rank + 1) * Ks])
    # replace the queue by the golden one so the reference's own numbers apply
    moco.memory.copy_(torch.from_numpy(g["memory_before"])[rank * Ks:(rank + 1) * Ks])
    ref_mem = torch.from_numpy(g["memory_before"]).clone()
    ref_idx = 0
    crit = NCESoftmaxLoss()
    B_loc = 8 // world
    for st in range(int(g["steps"])):
        q_all, k_all = torch.from_numpy(g[f"q{st}"]), torch.from_numpy(g[f"k{st}"])
        q = q_all[rank * B_loc:(rank + 1) * B_loc].clone().requires_grad_(True)
        k = k_all[rank * B_loc:(rank + 1) * B_loc]
        out, labels = moco(q, k)                                   # local rows in, keys gathered inside
        loss = crit(out)
        loss.backward()
        # what ONE replica of the reference computes for its local rows against the full (replicated) queue
        o = oracle.infonce_step(q.detach(), k, ref_mem.clone(), 0, T)
        assert abs(float(loss) - float(o["loss"])) <= 1e-5 * abs(float(o["loss"])), (float(loss), float(o["loss"]))
        assert torch.allclose(q.grad, o["dq"], rtol=1e-4, atol=1e-8)
        assert torch.equal(out.rank.long(), o["rank"])
        assert labels.shape == (B_loc,) and out.shape == (B_loc, K + 1)
        ref_idx = oracle.enqueue(ref_mem, k_all, ref_idx)           # every replica enqueues the gathered keys
        assert moco.index == ref_idx == int(g[f"index_after{st}"])
        assert torch.equal(moco.memory, ref_mem[rank * Ks:(rank + 1) * Ks])     # each rank wrote exactly the slots it owns
    full = moco.gather_full_queue()
    assert torch.equal(full, torch.from_numpy(g["memory_after"]))
    # wrap-around with pre-gathered keys passed by the caller (all_k=...), pointer near the end of the ring
    moco.index = ref_idx = K - 5
    keys = torch.nn.functional.normalize(torch.randn(12, d, generator=torch.Generator().manual_seed(7)))
    q = torch.from_numpy(g["q0"])[rank * B_loc:(rank + 1) * B_loc].clone().requires_grad_(True)
    k = torch.from_numpy(g["k0"])[rank * B_loc:(rank + 1) * B_loc]
    with pytest.raises(ValueError):
        moco(q, k, all_k=keys[:3])                                  # gathered keys must cover every rank's rows
    # the trainer's all_k is in ShuffleBN (permuted) row order (train...:222, 231): it feeds the enqueue only -- the positives
    # are the local, un-shuffled k of every rank
    moco.index = ref_idx = 16
    ref_mem = moco.gather_full_queue().clone()
    q_all, k_all = torch.from_numpy(g["q1"]), torch.from_numpy(g["k1"])
    all_k = k_all[torch.randperm(8, generator=torch.Generator().manual_seed(3))]
    assert not torch.equal(all_k, k_all)
    q = q_all[rank * B_loc:(rank + 1) * B_loc].clone().requires_grad_(True)
    k = k_all[rank * B_loc:(rank + 1) * B_loc]
    out, labels = moco(q, k, all_k=all_k)
    loss = crit(out)
    loss.backward()
    o = oracle.infonce_step(q.detach(), k, ref_mem.clone(), 0, T)
    assert abs(float(loss) - float(o["loss"])) <= 1e-5 * abs(float(o["loss"])), (float(loss), float(o["loss"]))
    assert torch.allclose(q.grad, o["dq"], rtol=1e-4, atol=1e-6), (rank, float((q.grad - o["dq"]).abs().max()), float(o["dq"].abs().max()))
    assert torch.equal(out.rank.long(), o["rank"])
    ref_idx = oracle.enqueue(ref_mem, all_k, ref_idx)               # the rows every replica of the reference enqueues
    assert moco.index == ref_idx and torch.equal(moco.gather_full_queue(), ref_mem)
    # checkpoints: the collective full_state_dict() is the upstream format; loading a full queue keeps the owned slots
    moco.index = 77
    sd = moco.full_state_dict(include_pointer=True)
    assert list(sd.keys()) == ["memory", "index"] and sd["memory"].dtype == torch.float32 and sd["memory"].shape == (K, d)
    assert torch.equal(sd["memory"], moco.gather_full_queue())
    fresh = ShardedRGBMoCo(d, K=K, T=T, compute=OracleCompute())
    fresh.load_state_dict(sd)
    assert torch.equal(fresh.memory, moco.memory) and fresh.index == 77
    upstream = {"memory": sd["memory"].clone()}                       # what an unmodified RGBMoCo checkpoint holds
    fresh2 = ShardedRGBMoCo(d, K=K, T=T, compute=OracleCompute())
    fresh2.load_state_dict(upstream)
    assert torch.equal(fresh2.memory, sd["memory"][rank * Ks:(rank + 1) * Ks]) and fresh2.index == 0
    local = moco.state_dict()                                          # non-collective: this rank's shard + its slot range
    assert list(local.keys()) == ["memory", "shard_begin", "shard_world"] and torch.equal(local["memory"], moco.memory.float())
    assert int(local["shard_begin"]) == rank * Ks and int(local["shard_world"]) == world
    fresh2.load_state_dict(dict(local))
    assert torch.equal(fresh2.memory, moco.memory)
    # what an unmodified trainer would do: rank 0 saves its state_dict(), every rank resumes from it -> loud failure on rank != 0
    everyone = [None] * world
    dist.all_gather_object(everyone, {k_: v.clone() for k_, v in local.items()})
    if rank != 0:
        with pytest.raises((ValueError, RuntimeError)):
            fresh2.load_state_dict(dict(everyone[0]))
        with pytest.raises((ValueError, RuntimeError)):
            fresh2.load_state_dict({"memory": everyone[0]["memory"]})   # a bare shard names no slots
    with pytest.raises((ValueError, RuntimeError)):
        fresh2.load_state_dict({"memory": torch.zeros(K + 8, d)})
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    dist.destroy_process_group()


def test_sharded_moco_world2(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))


def _shuffle_worker(rank, world, port, out_dir):
    for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gca_b200.dist import ShuffleBN, shuffle_plan
    from oracle.shuffle import shuffle_bn_all_ranks
    bsz = 6
    gen = torch.Generator().manual_seed(5)
    xs = [torch.randn(bsz, 3, 2, 4, 4, generator=gen) for _ in range(world)]          # every rank can rebuild all clips
    w = torch.randn(3 * 2 * 4 * 4, 16, generator=gen)

    def encoder(x):                                   # batch-dependent on purpose (BN-like): the shuffle must be exact
        f = x.reshape(x.shape[0], -1) @ w
        return torch.nn.functional.normalize(f - f.mean(0, keepdim=True))

    sbn = ShuffleBN()
    for it in range(3):
        torch.manual_seed(40 + it + 1000 * rank)      # ranks draw different permutations; rank 0's must win
        k, all_k = sbn(xs[rank], encoder)
        torch.manual_seed(40 + it)
        ids0 = torch.randperm(bsz * world)
        assert torch.equal(sbn.last_shuffle_ids, ids0)
        ref_k, ref_all_k, ref_this = shuffle_bn_all_ranks(xs, encoder, ids0)
        assert torch.equal(sbn.exchange(xs[rank], ids0), ref_this[rank])
        assert torch.equal(all_k, ref_all_k)
        assert torch.equal(k, ref_k[rank])
        rows, sc, rc, place = shuffle_plan(ids0, bsz, world, rank)
        assert sum(sc) == bsz and sum(rc) == bsz and sorted(place.tolist()) == list(range(bsz))
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_shuffle_bn_matches_reference_indexing(tmp_path, world):
    port = 31500 + (os.getpid() % 2000) + world
    mp.spawn(_shuffle_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))


def _make_encoder(seed):
    """Same construction as oracle/gen_golden_dist.py::make_encoder (batch-dependent: BatchNorm in train mode)."""
    g = torch.Generator().manual_seed(seed)
    net = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(24, 16), torch.nn.BatchNorm1d(16), torch.nn.ReLU(),
                              torch.nn.Linear(16, 8))
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * 0.5)
    net.train()
    return net


def _shuffle_fixture_worker(rank, world, port, out_dir):
    for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from gca_b200.dist import ShuffleBN
    g = dict(np.load(os.path.join(GOLDEN, "shuffle_bn_w2.npz")))
    assert int(g["world"]) == world
    enc = _make_encoder(int(g["encoder_seed"]))
    sbn = ShuffleBN()
    for it in range(int(g["iters"])):
        x = torch.from_numpy(g["r%d_x%d" % (rank, it)])
        torch.manual_seed(7 + it + 50 * rank)                         # the seeds the fixture generator used
        k, all_k = sbn(x, enc)
        assert torch.equal(all_k, torch.from_numpy(g["r%d_all_k%d" % (rank, it)])), ("all_k", it)
        assert torch.equal(k, torch.from_numpy(g["r%d_k%d" % (rank, it)])), ("k", it)
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    dist.destroy_process_group()


def test_shuffle_bn_against_reference_fixture(tmp_path):
    """tests/golden/shuffle_bn_w2.npz holds what the reference's own Trainer._shuffle_bn returned under gloo at world size 2
    (oracle/gen_golden_dist.py executes the method bodies from the reference file): the all-to-all version must return
    the same k and all_k bit for bit -- same permutation draw, same rows through the batch-dependent encoder."""
    world = 2
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_shuffle_fixture_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))
