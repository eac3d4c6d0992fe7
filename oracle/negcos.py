"""Oracle: SimSiam negative-cosine loss `D`.  TEST INFRASTRUCTURE ONLY.

Restates lib/memory/criterion.py:47-62 (fun_type 'v2'; duplicate at
lib/modeling/graph_wrappers.py:93-108): `-cosine_similarity(p, z.detach(), dim=-1).mean()`.
ATen's cosine_similarity (eps = 1e-8) computes  sum(p*z) / sqrt(clamp_min(|p|^2 * |z|^2, eps^2)).
"""
import torch
import torch.nn.functional as F

COS_EPS = 1e-8


def neg_cosine(p, z):
    """Scalar loss; gradient flows to `p` only (z is detached, criterion.py:60)."""
    return -F.cosine_similarity(p, z.detach(), dim=-1, eps=COS_EPS).mean()


def neg_cosine_grad(p, z, grad_out=1.0):
    """Closed-form dL/dp for rows whose norm product is above eps:
    d/dp [-(p.z)/(|p||z|)] / B = -( z/(|p||z|) - (p.z) p /(|p|^3 |z|) ) / B."""
    with torch.no_grad():
        B = p.shape[0]
        pn = p.norm(dim=-1, keepdim=True)
        zn = z.norm(dim=-1, keepdim=True)
        dot = (p * z).sum(dim=-1, keepdim=True)
        return -(z / (pn * zn) - dot * p / (pn ** 3 * zn)) * (grad_out / B)
