#!/usr/bin/env python
"""Golden fixtures for the trainer-side plumbing (tests/golden/shuffle_bn_w2.npz, ema.npz): EXECUTES THE REFERENCE's own
`Trainer._shuffle_bn`, `Trainer._global_gather` and `Trainer._momentum_update` (tools/train_video_contrast_dis.py:176-231).
The trainer module cannot be imported here (apex, yacs, tensorboardX ... are not installed), so the three method bodies are
taken from the reference file with `ast` at generation time and compiled as they stand -- nothing is copied into this
repository.  `_shuffle_bn` runs under torch.distributed (gloo, world size 2, one process per rank).  One non-invasive
monkeypatch: torch.Tensor.cuda -> identity (the reference hard-codes `.cuda()`).
TEST INFRASTRUCTURE ONLY; build container only (needs /root/reference).

usage:  python oracle/gen_golden_dist.py [--ref /root/reference] [--out tests/golden]
"""
import argparse
import ast
import os
import sys
import types

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
WANTED = ("_momentum_update", "_global_gather", "_shuffle_bn")


def reference_methods(ref):
    """{name: function} compiled from the reference's Trainer class body (decorators dropped: plain functions)."""
    path = os.path.join(ref, "tools", "train_video_contrast_dis.py")
    tree = ast.parse(open(path).read(), filename=path)
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef):
            for fn in node.body:
                if isinstance(fn, ast.FunctionDef) and fn.name in WANTED:
                    fn.decorator_list = []
                    mod = ast.Module(body=[fn], type_ignores=[])
                    ns = {"torch": torch, "dist": dist}
                    exec(compile(mod, path, "exec"), ns)
                    out[fn.name] = ns[fn.name]
    assert set(out) == set(WANTED), sorted(out)
    return out


def make_encoder(seed):
    """A batch-dependent momentum encoder (BatchNorm in train mode), so a wrong shuffle changes the keys."""
    g = torch.Generator().manual_seed(seed)
    net = nn.Sequential(nn.Flatten(), nn.Linear(24, 16), nn.BatchNorm1d(16), nn.ReLU(), nn.Linear(16, 8))
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * 0.5)
    net.train()
    return net


def shuffle_worker(rank, world, port, ref, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.Tensor.cuda = lambda self, *a, **k: self
    fns = reference_methods(ref)
    me = types.SimpleNamespace(
        args=types.SimpleNamespace(local_rank=rank, node_rank=0, ngpus_per_node=world),
        local_group=dist.group.WORLD,
        cfg=types.SimpleNamespace(CONTRAST=types.SimpleNamespace(JIGSAW=False)))
    me._global_gather = fns["_global_gather"]                  # a staticmethod upstream: called as self._global_gather(k)
    bsz = 6
    enc = make_encoder(3)
    res = {}
    for it in range(2):
        x = torch.randn(bsz, 2, 3, 2, 2, generator=torch.Generator().manual_seed(100 * it + rank))
        torch.manual_seed(7 + it + 50 * rank)                   # ranks draw different permutations; rank 0's wins
        k, all_k = fns["_shuffle_bn"](me, x, enc)
        res["x%d" % it], res["k%d" % it], res["all_k%d" % it] = x.numpy(), k.detach().numpy(), all_k.detach().numpy()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **res)
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(HERE), "tests", "golden"))
    args = ap.parse_args()
    torch.set_num_threads(1)
    import tempfile
    world = 2
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(shuffle_worker, args=(world, 29871, args.ref, tmp), nprocs=world, join=True)
        merged = {"world": np.int64(world), "encoder_seed": np.int64(3), "iters": np.int64(2)}
        for r in range(world):
            d = np.load(os.path.join(tmp, "rank%d.npz" % r))
            for key in d.files:
                merged["r%d_%s" % (r, key)] = d[key]
    np.savez_compressed(os.path.join(args.out, "shuffle_bn_w2.npz"), **merged)
    print("shuffle_bn_w2.npz:", sorted(merged)[:6], "...")

    # ---- EMA: the reference loop on a small model pair
    fns = reference_methods(args.ref)
    torch.manual_seed(21)
    make = lambda: nn.Sequential(nn.Conv3d(3, 5, 3), nn.BatchNorm3d(5), nn.Linear(5, 7), nn.Linear(7, 2, bias=False))
    model, ema = make(), make()
    before = [p.detach().clone().numpy() for p in ema.parameters()]
    src = [p.detach().clone().numpy() for p in model.parameters()]
    fns["_momentum_update"](model, ema, 0.999)
    fns["_momentum_update"](model, ema, 0.5)
    after = [p.detach().clone().numpy() for p in ema.parameters()]
    blob = {"n": np.int64(len(before)), "ms": np.array([0.999, 0.5])}
    for i, (b, s_, a) in enumerate(zip(before, src, after)):
        blob["before%d" % i], blob["src%d" % i], blob["after%d" % i] = b, s_, a
    np.savez_compressed(os.path.join(args.out, "ema.npz"), **blob)
    print("ema.npz: %d tensors" % len(before))


if __name__ == "__main__":
    main()
