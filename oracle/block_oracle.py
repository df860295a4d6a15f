"""CPU oracle (numpy, float64) for the frozen ViLT block ops this repo runs with its own kernels -- TEST INFRASTRUCTURE
ONLY (see oracle/__init__.py).

The reference does not implement these: they belong to the third-party backbone it instantiates
(``transformers.ViltModel``, reference src/modeling/vilt.py:19,127: ``ViltSelfAttention.forward`` -- scores = q k^T /
sqrt(d), softmax, context = probs v -- and ``ViltIntermediate.forward`` -- dense + exact (erf) GELU).  Pinned against
torch's own float64 ops and autograd in tests/test_block_oracle.py.
"""
import math

import numpy as np


def attention_forward(q, k, v, scale):
    """q, k, v: [B, S, H, D].  Returns (context [B, S, H, D], logsumexp of the scaled scores [B, H, S])."""
    q, k, v = (np.asarray(t, np.float64).transpose(0, 2, 1, 3) for t in (q, k, v))
    s = q @ k.transpose(0, 1, 3, 2) * scale
    m = s.max(-1, keepdims=True)
    e = np.exp(s - m)
    z = e.sum(-1, keepdims=True)
    return ((e / z) @ v).transpose(0, 2, 1, 3), (m + np.log(z))[..., 0]


def attention_backward(do, q, k, v, scale):
    """Gradients of ``attention_forward``'s context w.r.t. q, k, v for the upstream gradient ``do`` [B, S, H, D]."""
    q, k, v, do = (np.asarray(t, np.float64).transpose(0, 2, 1, 3) for t in (q, k, v, do))
    s = q @ k.transpose(0, 1, 3, 2) * scale
    p = np.exp(s - s.max(-1, keepdims=True))
    p /= p.sum(-1, keepdims=True)
    dv = p.transpose(0, 1, 3, 2) @ do
    dp = do @ v.transpose(0, 1, 3, 2)
    ds = p * (dp - (dp * p).sum(-1, keepdims=True))          # == p * (dp - rowsum(do * o))
    dq = ds @ k * scale
    dk = ds.transpose(0, 1, 3, 2) @ q * scale
    return tuple(t.transpose(0, 2, 1, 3) for t in (dq, dk, dv))


def gelu(x):
    """Exact GELU, x Phi(x) (HF ``GELUActivation`` / ``nn.GELU(approximate='none')``)."""
    x = np.asarray(x, np.float64)
    return 0.5 * x * (1.0 + np.vectorize(math.erf)(x / math.sqrt(2.0)))


def gelu_grad(x):
    x = np.asarray(x, np.float64)
    return 0.5 * (1.0 + np.vectorize(math.erf)(x / math.sqrt(2.0))) + x * np.exp(-0.5 * x * x) / math.sqrt(2.0 * math.pi)


def mlp_fc1_gelu(a, w, b):
    """(pre, act) = (a w^T + b, gelu(pre)) -- ViltIntermediate."""
    pre = np.asarray(a, np.float64) @ np.asarray(w, np.float64).T + np.asarray(b, np.float64)
    return pre, gelu(pre)


def mlp_fc2_dgelu(dy, w2, pre):
    """dpre = (dy w2) * gelu'(pre): the data gradient at ViltIntermediate's output pushed through the activation
    (w2 = ViltOutput.dense.weight, [out, hidden])."""
    return (np.asarray(dy, np.float64) @ np.asarray(w2, np.float64)) * gelu_grad(pre)
