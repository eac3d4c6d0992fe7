"""CPU restatement of the SimSiam projection / prediction MLPs (lib/modeling/project_head.py:36-76).  TEST INFRASTRUCTURE ONLY.

numpy fp64: Linear (x W^T + b), BatchNorm1d in training mode (biased batch variance normalises; running statistics move by
`momentum` with the unbiased variance, as nn.BatchNorm1d), ReLU; the backward of each block written out by hand."""
import numpy as np


def linear(x, W, b):
    """nn.Linear (project_head.py:39, 44, 49, 67, 71)."""
    return x @ W.T + (0.0 if b is None else b)


def bn1d_train(x, gamma, beta, eps=1e-5):
    """nn.BatchNorm1d forward in training mode (project_head.py:40, 45, 50, 68): returns y, (xhat, invstd), batch mean,
    unbiased batch variance (what the running statistics absorb)."""
    B = x.shape[0]
    mean = x.mean(0)
    var = ((x - mean) ** 2).sum(0) / B
    invstd = 1.0 / np.sqrt(var + eps)
    xhat = (x - mean) * invstd
    return xhat * gamma + beta, (xhat, invstd), mean, var * B / max(B - 1, 1)


def bn1d_train_backward(dy, cache, gamma):
    xhat, invstd = cache
    B = dy.shape[0]
    dbeta = dy.sum(0)
    dgamma = (dy * xhat).sum(0)
    dx = gamma * invstd * (dy - dbeta / B - xhat * dgamma / B)
    return dx, dgamma, dbeta


def block_forward(x, W, b, gamma, beta, relu, eps=1e-5):
    """One `Sequential(Linear, BatchNorm1d[, ReLU])` block; cache for `block_backward`."""
    z = linear(x, W, b)
    y, cache, mean, var_unb = bn1d_train(z, gamma, beta, eps)
    out = np.maximum(y, 0.0) if relu else y
    return out, dict(x=x, W=W, bn=cache, gamma=gamma, mask=(y > 0) if relu else None, mean=mean, var_unb=var_unb)


def block_backward(dout, c):
    dy = dout * c["mask"] if c["mask"] is not None else dout
    dz, dgamma, dbeta = bn1d_train_backward(dy, c["bn"], c["gamma"])
    return dz @ c["W"], dict(dW=dz.T @ c["x"], db=dz.sum(0), dgamma=dgamma, dbeta=dbeta)


def projection_mlp(x, p, eps=1e-5):
    """ProjectionMLP.forward (project_head.py:53-58) on a dict of parameters keyed like its state_dict()."""
    h1, c1 = block_forward(x, p["l1.0.weight"], p["l1.0.bias"], p["l1.1.weight"], p["l1.1.bias"], True, eps)
    h2, c2 = block_forward(h1, p["l2.0.weight"], p["l2.0.bias"], p["l2.1.weight"], p["l2.1.bias"], True, eps)
    h3, c3 = block_forward(h2, p["l3.0.weight"], p["l3.0.bias"], p["l3.1.weight"], p["l3.1.bias"], False, eps)
    return h3, (c1, c2, c3)


def projection_mlp_backward(dout, caches):
    c1, c2, c3 = caches
    d2, g3 = block_backward(dout, c3)
    d1, g2 = block_backward(d2, c2)
    dx, g1 = block_backward(d1, c1)
    return dx, (g1, g2, g3)


def prediction_mlp(x, p, eps=1e-5):
    """PredictionMLP.forward (project_head.py:73-77)."""
    h1, c1 = block_forward(x, p["l1.0.weight"], p["l1.0.bias"], p["l1.1.weight"], p["l1.1.bias"], True, eps)
    return linear(h1, p["l2.weight"], p["l2.bias"]), (c1, dict(h1=h1, W=p["l2.weight"]))


def prediction_mlp_backward(dout, caches):
    c1, c2 = caches
    g2 = dict(dW=dout.T @ c2["h1"], db=dout.sum(0))
    dx, g1 = block_backward(dout @ c2["W"], c1)
    return dx, (g1, g2)
