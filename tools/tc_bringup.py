#!/usr/bin/env python
"""Bring-up aid for the tcgen05 InfoNCE kernel: runs one small problem, prints per-quantity errors against a torch
fp64 restatement on bf16-rounded inputs.  With --sweep it re-runs itself under alternative UMMA descriptor knobs
(GCA_TC_DESC) to locate a descriptor mistake in one GPU session.  Not part of the product or the tests."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-graph-ssl_b200"))


def run_once(B, K, want_grad=True):
    import torch
    import torch.nn.functional as F
    from gca_b200 import functional as GF
    torch.manual_seed(0)
    mem = F.normalize(torch.randn(K, 128)).to(torch.bfloat16)
    q, k = F.normalize(torch.randn(B, 128)), F.normalize(torch.randn(B, 128))
    T = 0.07
    r = GF.infonce_forward(q.cuda(), k.cuda(), mem.cuda(), T, algo="tcgen05", want_grad=want_grad, materialize=True)
    torch.cuda.synchronize()
    rq = q.to(torch.bfloat16).double()
    pos = (q.double() * k.double()).sum(1) / T
    neg = rq @ mem.double().t() / T
    lse = torch.logsumexp(torch.cat([pos[:, None], neg], 1), 1)
    lg = r["logits"].double().cpu()
    e_logit = float((lg[:, 1:] - neg).abs().max())
    e_lse = float((r["lse"].double().cpu() - lse).abs().max())
    out = "B=%d K=%d  max|logit err|=%.3e  max|lse err|=%.3e" % (B, K, e_logit, e_lse)
    if want_grad:
        dq_ref = ((torch.exp(pos - lse) - 1)[:, None] * k.double() + torch.exp(neg - lse[:, None]) @ mem.double()) / (T * B)
        e_dq = float((r["dq_unit"].double().cpu() - dq_ref).abs().max() / dq_ref.abs().max())
        out += "  rel dq err=%.3e" % e_dq
    if e_logit > 1e-2:
        # where is it wrong? print a small corner of got / expected
        out += "\n got[0,:8]=%s\n exp[0,:8]=%s" % (lg[0, 1:9].tolist(), neg[0, :8].tolist())
        out += "\n got[1,:4]=%s\n exp[1,:4]=%s" % (lg[1, 1:5].tolist() if B > 1 else [], neg[1, :4].tolist() if B > 1 else [])
    print(out, flush=True)
    return e_logit, e_lse


if __name__ == "__main__":
    if False and "--sweep" in sys.argv:   # descriptor knobs were removed once the layouts were validated on hardware
        variants = ["16,1024,32,16384,1024,2048",      # canonical
                    "16,1024,32,1024,16384,2048",      # GEMM2 LBO/SBO swapped
                    "1,1024,32,16384,1024,2048",
                    "16,1024,32,16384,1024,256",       # GEMM2 k-step = 2 x 128-byte rows
                    "16,1024,32,128,1024,2048",
                    "1024,1024,32,16384,1024,2048"]
        for v in variants:
            env = dict(os.environ, GCA_TC_DESC=v)
            print("=== GCA_TC_DESC=%s" % v, flush=True)
            p = subprocess.run([sys.executable, __file__, "128", "128"], env=env, capture_output=True, text=True, timeout=120)
            print(p.stdout[-1500:], p.stderr[-800:], flush=True)
        sys.exit(0)
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    run_once(B, K, want_grad=False)
    run_once(B, K, want_grad=True)
