#!/usr/bin/env python
"""tests/golden/r3d18_shapes.json: the parameter shapes and the feature-map shape of the reference's pre-train backbone,
`resnet18(sample_size=112, sample_duration=16)` (lib/modeling/backbone/backbone_3d/resnet.py:108-222), taken by instantiating
the reference's own class on the CPU.  tools/pretrain_step.py's encoder (the workload around the head in the clips/s figure)
is checked against this list.  TEST INFRASTRUCTURE ONLY; needs /root/reference, the .json is committed."""
import importlib.util
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def main(ref="/root/reference"):
    spec = importlib.util.spec_from_file_location("ref_resnet", os.path.join(ref, "lib/modeling/backbone/backbone_3d/resnet.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    net = m.resnet18(sample_size=112, sample_duration=16)
    shapes = sorted(tuple(p.shape) for n, p in net.named_parameters() if not n.startswith("fc."))
    with torch.no_grad():
        f = net.maxpool(net.relu(net.bn1(net.conv1(torch.randn(1, 3, 16, 112, 112)))))
        f = net.layer4(net.layer3(net.layer2(net.layer1(f))))
    out = {"source": "lib/modeling/backbone/backbone_3d/resnet.py: resnet18(sample_size=112, sample_duration=16), parameters without fc",
           "sorted_param_shapes": [list(s) for s in shapes],
           "n_params": int(sum(p.numel() for n, p in net.named_parameters() if not n.startswith("fc."))),
           "feature_map_shape_for_1x3x16x112x112": list(f.shape)}
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "r3d18_shapes.json")
    json.dump(out, open(path, "w"))
    print("wrote", path, out["n_params"], "parameters in", len(shapes), "tensors")


if __name__ == "__main__":
    main(*sys.argv[1:])
