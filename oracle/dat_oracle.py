"""Oracle: the DAT bottleneck operator (reference src/modeling/models/adapter.py).  TEST ONLY.

Weights use the reference's nn.Linear layout: down.weight [r, d], down.bias [r], up.weight [d, r],
up.bias [d] (adapter.py:35,41).  A "branch" is the tuple (down_w, down_b, up_w, up_b).
"""
from __future__ import annotations

import math

import numpy as np


def _act(p, act):
    if act == "relu":                      # adapter.py:24  self.actv = nn.ReLU()
        return np.maximum(p, 0.0)
    if act == "gelu":                      # opt-in, erf form (torch.nn.GELU default)
        erf = np.vectorize(math.erf)
        return 0.5 * p * (1.0 + erf(p / math.sqrt(2.0)))
    raise ValueError(act)


def _act_grad(p, act):
    if act == "relu":
        return (p > 0.0).astype(p.dtype)
    if act == "gelu":
        erf = np.vectorize(math.erf)
        return 0.5 * (1.0 + erf(p / math.sqrt(2.0))) + p * np.exp(-0.5 * p * p) / math.sqrt(2.0 * math.pi)
    raise ValueError(act)


def adapter_forward(hidden_states, input_tensor, branches, gating, scaling=1.0, act="relu",
                    dtype=np.float64):
    """adapter.py:124-163.

    single mode (:125-131): ``branches`` holds the one active branch; out = input + up(act(down(h))).
    gating mode (:133-146 / :148-162): two branches, fixed 0.5/0.5 weights (:144), times
    ``self.scaling`` (= 1.0, adapter.py:25).
    """
    h = np.asarray(hidden_states, dtype=dtype)
    res = np.asarray(input_tensor, dtype=dtype)
    ups = []
    for (dw, db, uw, ub) in branches:
        down = h @ np.asarray(dw, dtype).T + np.asarray(db, dtype)       # :127 / :137
        down = _act(down, act)                                           # :128 / :138
        ups.append(down @ np.asarray(uw, dtype).T + np.asarray(ub, dtype))  # :129 / :140
    if not gating:
        assert len(branches) == 1
        return res + ups[0]                                              # :131
    assert len(branches) == 2
    agg = 0.5 * ups[0]                                                   # get_agg_out :118-122
    agg = agg + 0.5 * ups[1]
    return res + agg * scaling                                           # :146


def adapter_backward(hidden_states, grad_out, branches, gating, residual_is_input, scaling=1.0,
                     act="relu", dtype=np.float64):
    """Analytic gradient of adapter_forward (what torch autograd computes for adapter.py:124-163).

    Returns (d_hidden, [per-branch (d_down_w, d_down_b, d_up_w, d_up_b)]).  When
    ``residual_is_input`` the same tensor is passed as hidden_states and input_tensor
    (adaptered_output.py:78, vit.py:107), so d_hidden also carries grad_out.
    """
    h = np.asarray(hidden_states, dtype=dtype)
    g = np.asarray(grad_out, dtype=dtype)
    h2 = h.reshape(-1, h.shape[-1])
    g2 = g.reshape(-1, g.shape[-1])
    s = (0.5 * scaling) if gating else 1.0
    dh = np.zeros_like(h2)
    grads = []
    for (dw, db, uw, ub) in branches:
        dw = np.asarray(dw, dtype); db = np.asarray(db, dtype); uw = np.asarray(uw, dtype)
        p = h2 @ dw.T + db
        hid = _act(p, act)
        d_up = s * g2
        d_up_w = d_up.T @ hid
        d_up_b = d_up.sum(0)
        d_hid = d_up @ uw
        d_p = d_hid * _act_grad(p, act)
        d_down_w = d_p.T @ h2
        d_down_b = d_p.sum(0)
        dh = dh + d_p @ dw
        grads.append((d_down_w, d_down_b, d_up_w, d_up_b))
    if residual_is_input:
        dh = dh + g2
    return dh.reshape(h.shape), grads


def layer_norm(x, weight, bias, eps):
    x = np.asarray(x, dtype=np.float64)
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * weight + bias


def adapter_layer_forward_bert(hidden_states, input_tensor, ln_weight, ln_bias, ln_eps, branches,
                               gating, scaling=1.0, act="relu"):
    """adapter.py:97-116: pre_forward (residual = ffn output, LN(ffn + x)), forward, post_forward
    (the SAME LayerNorm applied again to adapter_out + x)."""
    residual = np.asarray(hidden_states, np.float64)                               # :104
    h = layer_norm(residual + np.asarray(input_tensor, np.float64), ln_weight, ln_bias, ln_eps)  # :106
    h = adapter_forward(h, residual, branches, gating, scaling, act)               # :99
    return layer_norm(h + np.asarray(input_tensor, np.float64), ln_weight, ln_bias, ln_eps)      # :113


def pack_branches(branches, dtype=np.float32):
    """The concatenated operands the CUDA kernels consume (include/feddat_b200.h,
    feddat_pack_weights): Wd_cat [nR, d], bd_cat [nR], Wu_cat [d, nR], bu_cat [d] = sum of up biases."""
    wd = np.concatenate([np.asarray(b[0], dtype) for b in branches], axis=0)
    bd = np.concatenate([np.asarray(b[1], dtype) for b in branches], axis=0)
    wu = np.concatenate([np.asarray(b[2], dtype) for b in branches], axis=1)
    bu = np.sum([np.asarray(b[3], dtype) for b in branches], axis=0)
    return wd, bd, wu, bu
