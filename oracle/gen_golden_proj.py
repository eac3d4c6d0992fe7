#!/usr/bin/env python
"""Golden fixture for the projection-tail fusion (tests/golden/proj_tail.npz): RUNS THE REFERENCE's own `Normalize`
(lib/modeling/project_head.py:4-10), `RGBMoCo` and `NCESoftmaxLoss` on seeded un-normalised projections, with autograd
through the normalisation, and checks `oracle.infonce.head_from_projections` against them.  TEST INFRASTRUCTURE ONLY;
build container only (needs /root/reference).

usage:  python oracle/gen_golden_proj.py [--ref /root/reference] [--out tests/golden]
"""
import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.dont_write_bytecode = True


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(HERE), "tests", "golden"))
    args = ap.parse_args()
    sys.path.insert(0, args.ref)
    torch.Tensor.cuda = lambda self, *a, **k: self          # `.cuda()` is hard-coded in mem_moco.py
    from lib.memory.mem_moco import RGBMoCo
    from lib.memory.criterion import NCESoftmaxLoss
    from lib.modeling.project_head import Normalize
    import oracle
    from oracle.infonce import head_from_projections
    torch.set_num_threads(1)
    B, K, d, T = 48, 512, 128, 0.07
    torch.manual_seed(11)
    moco = RGBMoCo(d, K=K, T=T)
    mem_before = moco.memory.clone()
    zq = (torch.randn(B, d) * 2.5).requires_grad_(True)
    zk = torch.randn(B, d) * 0.3
    norm = Normalize(2)
    out, labels = moco(norm(zq), norm(zk))
    loss = NCESoftmaxLoss()(out)
    loss.backward()
    o = head_from_projections(zq.detach(), zk, mem_before.clone(), 0, T)
    assert abs(float(loss.detach()) - float(o["loss"])) <= 2e-6 * abs(float(loss.detach()))
    assert float((zq.grad - o["dz"]).abs().max()) <= 1e-4 * float(zq.grad.abs().max())
    assert torch.equal(moco.memory[:B], o["k_hat"]) and moco.index == B
    np.savez_compressed(os.path.join(args.out, "proj_tail.npz"), zq=zq.detach().numpy(), zk=zk.numpy(),
                        memory_before=mem_before.numpy(), enqueued_rows=moco.memory[:B].numpy(), T=np.float32(T),
                        loss=np.float64(float(loss.detach())), dz=zq.grad.numpy(), index_after=np.int64(moco.index))
    print("proj_tail.npz: loss %.10f, |dz|_1 %.6f" % (float(loss.detach()), float(zq.grad.abs().sum())))


if __name__ == "__main__":
    main()
