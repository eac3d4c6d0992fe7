#!/usr/bin/env python
"""Golden fixture for the SimSiam MLPs (tests/golden/simsiam_mlp.npz): RUNS THE REFERENCE's own `ProjectionMLP` and
`PredictionMLP` (lib/modeling/project_head.py:36-76) forward + autograd backward in training mode on seeded inputs, records
parameters, outputs, every gradient and the running statistics after the step, and checks `oracle.mlp` against them.
TEST INFRASTRUCTURE ONLY; build container only (needs /root/reference).

usage:  python oracle/gen_golden_mlp.py [--ref /root/reference] [--out tests/golden]
"""
import argparse
import importlib.util
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
    # project_head.py is self-contained; lib.modeling's package __init__ pulls the backbones in, so load the file itself
    spec = importlib.util.spec_from_file_location("ref_project_head", os.path.join(args.ref, "lib", "modeling", "project_head.py"))
    ph = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ph)
    from oracle import mlp as om
    torch.set_num_threads(1)
    B, in_dim, hid = 24, 40, 64
    out = {}
    torch.manual_seed(21)
    proj = ph.ProjectionMLP(in_dim, hid, hid).train()
    pred = ph.PredictionMLP(hid, hid // 2, hid).train()
    for m in (proj, pred):                                  # non-trivial affine parameters and running statistics
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm1d):
                mod.weight.data.uniform_(0.5, 1.5)
                mod.bias.data.uniform_(-0.3, 0.3)
                mod.running_mean.uniform_(-0.2, 0.2)
                mod.running_var.uniform_(0.5, 2.0)
    for name, m, x in (("proj", proj, torch.randn(B, in_dim)), ("pred", pred, torch.randn(B, hid))):
        sd0 = {k: v.clone() for k, v in m.state_dict().items()}
        x = x.requires_grad_(True)
        y = m(x)
        w = torch.randn_like(y)
        (y * w).sum().backward()
        p64 = {k: v.double().numpy() for k, v in sd0.items()}
        if name == "proj":
            yo, caches = om.projection_mlp(x.detach().double().numpy(), p64)
            dxo, grads = om.projection_mlp_backward(w.double().numpy(), caches)
        else:
            yo, caches = om.prediction_mlp(x.detach().double().numpy(), p64)
            dxo, grads = om.prediction_mlp_backward(w.double().numpy(), caches)
        assert np.abs(yo - y.detach().numpy()).max() <= 2e-5 * np.abs(yo).max(), name
        assert np.abs(dxo - x.grad.numpy()).max() <= 1e-4 * np.abs(dxo).max(), name
        g1 = grads[0]
        assert np.abs(g1["dW"] - m.l1[0].weight.grad.numpy()).max() <= 1e-4 * np.abs(g1["dW"]).max(), name
        assert np.abs(g1["dgamma"] - m.l1[1].weight.grad.numpy()).max() <= 1e-4 * max(np.abs(g1["dgamma"]).max(), 1e-3), name
        c1 = caches[0]
        rm = 0.9 * sd0["l1.1.running_mean"].numpy() + 0.1 * c1["mean"]
        rv = 0.9 * sd0["l1.1.running_var"].numpy() + 0.1 * c1["var_unb"]
        assert np.abs(rm - m.l1[1].running_mean.numpy()).max() <= 1e-5 and np.abs(rv - m.l1[1].running_var.numpy()).max() <= 1e-5
        out[name + ".x"] = x.detach().numpy(); out[name + ".w"] = w.numpy(); out[name + ".y"] = y.detach().numpy()
        out[name + ".dx"] = x.grad.numpy()
        for k, v in sd0.items():
            out[name + ".before." + k] = v.numpy()
        for k, v in m.state_dict().items():
            out[name + ".after." + k] = v.numpy()
        for k, p in m.named_parameters():
            out[name + ".grad." + k] = p.grad.numpy()
        print("%s: |y|_1 %.6f  |dx|_1 %.6f" % (name, float(y.detach().abs().sum()), float(x.grad.abs().sum())))
    np.savez_compressed(os.path.join(args.out, "simsiam_mlp.npz"), **out)


if __name__ == "__main__":
    main()
