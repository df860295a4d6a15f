"""The numpy oracle of the frozen block ops (oracle/block_oracle.py) against torch's float64 ops and autograd (CPU)."""
import numpy as np
import torch

import oracle


def test_attention_oracle_matches_torch_autograd():
    g = torch.Generator().manual_seed(0)
    B, S, H, D = 2, 37, 3, 16
    q, k, v, do = (torch.randn(B, S, H, D, generator=g, dtype=torch.float64) for _ in range(4))
    qt, kt, vt = (t.permute(0, 2, 1, 3).clone().requires_grad_(True) for t in (q, k, v))
    out = torch.nn.functional.scaled_dot_product_attention(qt, kt, vt, scale=0.25)
    out.backward(do.permute(0, 2, 1, 3))
    o, lse = oracle.attention_forward(q.numpy(), k.numpy(), v.numpy(), 0.25)
    dq, dk, dv = oracle.attention_backward(do.numpy(), q.numpy(), k.numpy(), v.numpy(), 0.25)
    np.testing.assert_allclose(o, out.detach().permute(0, 2, 1, 3).numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(lse, torch.logsumexp(qt @ kt.transpose(-1, -2) * 0.25, -1).detach().numpy(), rtol=1e-10)
    for got, want in ((dq, qt.grad), (dk, kt.grad), (dv, vt.grad)):
        np.testing.assert_allclose(got, want.permute(0, 2, 1, 3).numpy(), rtol=1e-9, atol=1e-12)


def test_mlp_oracle_matches_torch_autograd():
    g = torch.Generator().manual_seed(1)
    a = torch.randn(13, 24, generator=g, dtype=torch.float64)
    w = torch.randn(40, 24, generator=g, dtype=torch.float64) * 0.3
    b = torch.randn(40, generator=g, dtype=torch.float64)
    w2 = torch.randn(24, 40, generator=g, dtype=torch.float64) * 0.3
    dy = torch.randn(13, 24, generator=g, dtype=torch.float64)
    pre_t = (a @ w.T + b).requires_grad_(True)
    act_t = torch.nn.functional.gelu(pre_t)
    (act_t @ w2.T).backward(dy)
    pre, act = oracle.mlp_fc1_gelu(a.numpy(), w.numpy(), b.numpy())
    np.testing.assert_allclose(pre, pre_t.detach().numpy(), rtol=1e-12)
    np.testing.assert_allclose(act, act_t.detach().numpy(), rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(oracle.mlp_fc2_dgelu(dy.numpy(), w2.numpy(), pre), pre_t.grad.numpy(), rtol=1e-9, atol=1e-13)
