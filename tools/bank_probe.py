#!/usr/bin/env python
"""Instance bank at the size of the reference's bank mode (bsz 256, K = 16384 negatives, d = 128, n_data = 240k clips):
gca_bank_logits + gca_bank_dx + gca_bank_update against the reference's op sequence in eager PyTorch on the same GPU
(index_select -> bmm -> div, autograd, momentum update).  Device time per step, L2 flushed before every step."""
import json, os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-graph-ssl_b200"))
from gca_b200 import functional as GF

B, K, d, n_data, T, m = 256, 16384, 128, 240000, 0.07, 0.5
torch.manual_seed(0)
bank = F.normalize(torch.randn(n_data, d, device="cuda"))
x = F.normalize(torch.randn(B, d, device="cuda"))
y = torch.randperm(n_data, device="cuda")[:B]
idx = torch.randint(0, n_data, (B, K + 1), device="cuda")
idx[:, 0] = y
w = torch.randn(B, K + 1, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def ours():
    xg = x.clone().requires_grad_(True)
    lg = GF.bank_logits(xg, bank, idx, T)
    lg.backward(w)
    GF.bank_update_(bank, x, y, m)
    return xg.grad

def eager():
    xg = x.clone().requires_grad_(True)
    wt = torch.index_select(bank, 0, idx.view(-1)).view(B, K + 1, d)
    lg = torch.bmm(wt, xg.unsqueeze(2)).div(T).squeeze(2)
    lg.backward(w)
    with torch.no_grad():
        wp = torch.index_select(bank, 0, y)
        wp.mul_(m).add_(torch.mul(x, 1 - m))
        bank.index_copy_(0, y, F.normalize(wp))
    return xg.grad

def timeit(fn, n=8, warm=2):
    ts = []
    for i in range(n + warm):
        flush.fill_(i & 1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]

bank0 = bank.clone()
g1 = ours().clone()
bank.copy_(bank0)                                              # both variants score the same bank
g2 = eager().clone()
err = float((g1 - g2).abs().max() / g2.abs().max())
t_ours, t_eager = timeit(ours), timeit(eager)
gather_bytes = 2.0 * B * (K + 1) * d * 4                       # the sampled rows once per direction
print(json.dumps({"what": "instance bank step (logits + dx + update), bsz 256, K 16384, d 128, n_data 240000",
                  "ms": round(t_ours, 4), "eager_torch_ms": round(t_eager, 4), "gathered_GBps": round(gather_bytes / t_ours / 1e6, 1),
                  "dx_rel_err_vs_eager": err}))
