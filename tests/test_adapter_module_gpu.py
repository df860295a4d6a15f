"""The ``Adapter`` MODULE (autograd node included) against the oracle at the wide ranks of BASELINE
configs[2] (ALBEF, r = 256) and configs[4] (rank sweep up to r = 512): bottlenecks wider than one launch
covers (R = r or 2r > 256) run as several segment launches whose partial results the autograd node has to
stitch (``Adapter._segments``, ``_DatFunction.backward``).  Forward, dX, d(residual) and every trainable
parameter gradient, single and gating mode, residual == input (ViLT / ViT sites) and residual != input
(BERT sites, adapter.py:97-116).  GPU only."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

BF16_TOL = 1e-2


def relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def bf16(x):
    return torch.from_numpy(np.asarray(x, np.float32)).to(torch.bfloat16).float().numpy()


@pytest.mark.parametrize("same_residual", [True, False], ids=["res_is_x", "res_not_x"])
@pytest.mark.parametrize("gating", [False, True], ids=["single", "gating"])
@pytest.mark.parametrize("r", [48, 128, 256, 512])
def test_adapter_module_matches_oracle(r, gating, same_residual):
    from feddat_b200.modeling.adapter import Adapter
    M = 333                                              # 2 full tiles + a ragged one
    rng = np.random.default_rng(r * 13 + 2 * gating + same_residual)
    a = Adapter(names=["adapter_0", "adapter_1", "adapter_2"], device="cuda", rank=r)
    with torch.no_grad():
        for n, p in a.named_parameters():
            std = 0.1 if p.dim() == 1 else 0.05
            p.copy_(torch.from_numpy((rng.standard_normal(tuple(p.shape)) * std).astype(np.float32)))
    x = rng.standard_normal((3, 111, 768)).astype(np.float32)
    res = x if same_residual else rng.standard_normal((3, 111, 768)).astype(np.float32)
    g = rng.standard_normal((3, 111, 768)).astype(np.float32)
    assert x.shape[0] * x.shape[1] == M

    if gating:
        a.activate_gating()
        a.set_active_adapter("adapter_0")                # adapter_0 trains, adapter_2 is frozen
        names, train = ("adapter_0", "adapter_2"), "adapter_0"
    else:
        a.deactivate_gating()
        a.set_active_adapter("adapter_1")
        names, train = ("adapter_1",), "adapter_1"

    xd = torch.from_numpy(x).cuda().to(torch.bfloat16).requires_grad_(True)
    if same_residual:
        y = a(xd, xd)
        rd = None
    else:
        rd = torch.from_numpy(res).cuda().to(torch.bfloat16).requires_grad_(True)
        y = a(xd, rd)
    y.backward(torch.from_numpy(g).cuda().to(torch.bfloat16))
    torch.cuda.synchronize()

    def branch(n):
        d, u = getattr(a, f"{n}_down"), getattr(a, f"{n}_up")
        return (bf16(d.weight.detach().cpu().numpy()), d.bias.detach().cpu().numpy(),
                bf16(u.weight.detach().cpu().numpy()), u.bias.detach().cpu().numpy())

    brs = [branch(n) for n in names]
    xr, rr, gr = bf16(x.reshape(M, 768)), bf16(res.reshape(M, 768)), bf16(g.reshape(M, 768))
    y_or = oracle.adapter_forward(xr, rr, brs, gating)
    # one launch rounds the adapter output to bf16 once (6e-3 bar of tests/test_kernels_gpu.py); a bottleneck
    # split into n segment launches rounds n partial sums (measured 6.4e-3 at R = 1024): the bf16 bar 1e-2
    n_seg = -(-(len(names) * r) // 256)
    assert relerr(y.detach().float().cpu().numpy().reshape(M, 768), y_or) < (6e-3 if n_seg == 1 else BF16_TOL)
    dx_or, grads_or = oracle.adapter_backward(xr, gr, brs, gating, residual_is_input=same_residual)
    # rows holding a pre-activation within rounding noise of 0 may take either relu' value
    # (tests/test_kernels_gpu.py::test_dat_fwd_bwd_full_size_vs_oracle explains the exemption)
    near = np.zeros(M, bool)
    for (dw, db, _, _) in brs:
        near |= (np.abs(xr.astype(np.float64) @ dw.T.astype(np.float64) + db) < 2e-5).any(axis=1)
    assert near.sum() <= 8
    dx = xd.grad.float().cpu().numpy().reshape(M, 768)
    row_err = np.abs(dx - dx_or).max(axis=1) / np.abs(dx_or).max()
    assert set(np.nonzero(row_err >= BF16_TOL)[0].tolist()) <= set(np.nonzero(near)[0].tolist())
    if not same_residual:
        assert torch.equal(rd.grad, torch.from_numpy(g).cuda().to(torch.bfloat16))      # d(residual) = dY
    ti = names.index(train)
    d, u = getattr(a, f"{train}_down"), getattr(a, f"{train}_up")
    for got, want, what in zip((d.weight.grad, d.bias.grad, u.weight.grad, u.bias.grad), grads_or[ti],
                               ("down.weight", "down.bias", "up.weight", "up.bias")):
        assert got is not None, what
        assert relerr(got.cpu().numpy(), want) < 2 * BF16_TOL, what
    for n in names:
        if n != train:
            assert all(p.grad is None for p in getattr(a, f"{n}_down").parameters())


@pytest.mark.parametrize("r", [256, 512])
def test_bert_site_wrapper_wide_rank(r):
    """adapter_layer_forward_bert (adapter.py:97-116) at the ALBEF rank: LN(ffn + x) -> adapter(residual = ffn)
    -> LN(. + x), against the oracle's restatement evaluated with the kernel's bf16 operands."""
    from feddat_b200.modeling.adapter import Adapter
    rng = np.random.default_rng(r)
    a = Adapter(names=["adapter_0", "adapter_1", "adapter_2"], device="cuda", rank=r)
    with torch.no_grad():
        for n, p in a.named_parameters():
            p.copy_(torch.from_numpy((rng.standard_normal(tuple(p.shape)) * (0.1 if p.dim() == 1 else 0.05)).astype(np.float32)))
    a.activate_gating(); a.set_active_adapter("adapter_0")
    ln = torch.nn.LayerNorm(768, eps=1e-12).cuda()
    with torch.no_grad():
        ln.weight.copy_(1 + 0.1 * torch.randn(768, device="cuda")); ln.bias.copy_(0.1 * torch.randn(768, device="cuda"))
    ffn = rng.standard_normal((2, 25, 768)).astype(np.float32)
    x = rng.standard_normal((2, 25, 768)).astype(np.float32)
    # fp32 activations in (LayerNorm in fp32, as the reference under autocast); the operator rounds its own
    # input and residual to bf16
    out = a.adapter_layer_forward_bert(torch.from_numpy(ffn).cuda(), torch.from_numpy(x).cuda(), ln)

    def branch(n):
        d, u = getattr(a, f"{n}_down"), getattr(a, f"{n}_up")
        return (bf16(d.weight.detach().cpu().numpy()), d.bias.detach().cpu().numpy(),
                bf16(u.weight.detach().cpu().numpy()), u.bias.detach().cpu().numpy())

    want = oracle.adapter_layer_forward_bert(ffn, x, ln.weight.detach().cpu().numpy().astype(np.float64),
                                             ln.bias.detach().cpu().numpy().astype(np.float64), 1e-12,
                                             [branch("adapter_0"), branch("adapter_2")], True)
    assert relerr(out.detach().float().cpu().numpy(), want) < 2e-2       # two LayerNorms amplify the bf16 rounding of their inputs


def test_dual_mode_deferred_weight_gradients():
    """Several sites in dual mode (TaskTrainer's batched schedule) chained like encoder blocks: with
    ``ops.deferred_wgrad()`` around the backward, the data gradient is bit-identical, every parameter gets the
    gradient of the plain autograd run (one launch over all sites, no row splits: summation order differs), the
    frozen adapter_2 gets none, and a second deferred backward ACCUMULATES into .grad as autograd would."""
    from feddat_b200 import ops
    from feddat_b200.modeling.adapter import Adapter
    g = torch.Generator(device="cuda").manual_seed(11)
    sites = []
    for _ in range(4):
        a = Adapter(names=["adapter_0", "adapter_1", "adapter_2"], device="cuda", rank=64)
        with torch.no_grad():
            for p in a.parameters():
                p.copy_(torch.randn(p.shape, device="cuda", generator=g) * (0.1 if p.dim() == 1 else 0.05))
        a.set_dual(True)
        sites.append(a)
    x0 = torch.randn(2 * 333, 768, device="cuda", generator=g).to(torch.bfloat16)
    dy = torch.randn(2 * 333, 768, device="cuda", generator=g).to(torch.bfloat16)

    def run(defer, times=1):
        for a in sites:
            a.zero_grad(set_to_none=True)
        for _ in range(times):
            x = x0.clone().requires_grad_(True)
            h = x
            for a in sites:
                h = a(h, h)
            if defer:
                with ops.deferred_wgrad():
                    h.backward(dy)
            else:
                h.backward(dy)
        torch.cuda.synchronize()
        return x.grad.clone(), {n: (None if p.grad is None else p.grad.clone())
                                for i, a in enumerate(sites) for n, p in ((f"{i}.{k}", v) for k, v in a.named_parameters())}

    dx_ref, ref = run(False)
    dx_def, got = run(True)
    assert torch.equal(dx_ref, dx_def)
    assert ref.keys() == got.keys()
    for n in ref:
        if "adapter_2" in n:
            assert ref[n] is None and got[n] is None
        else:
            assert got[n].shape == ref[n].shape and got[n].is_contiguous()
            assert relerr(got[n].cpu().numpy(), ref[n].cpu().numpy()) < 1e-5, n
    _, twice = run(True, times=2)
    for n in ref:
        if ref[n] is not None:
            assert relerr(twice[n].cpu().numpy(), 2 * got[n].cpu().numpy()) < 1e-6, n
    assert ops.deferred_queue() is None


@pytest.mark.parametrize("r,gating", [(64, False), (64, True), (256, False), (256, True), (512, True)])
def test_reference_order_deferred_weight_gradients(r, gating):
    """``_DatFunction`` (single / gating mode, residual != input, bottlenecks of one, two and -- r = 512 -- four segment
    launches) under ``ops.deferred_wgrad()``: same dX / d(residual) bit for bit, same parameter gradients as the plain
    autograd run.  At r = 512 a branch is cut across segments and the node must fall back to immediate launches."""
    from feddat_b200 import ops
    from feddat_b200.modeling.adapter import Adapter
    g = torch.Generator(device="cuda").manual_seed(r + gating)
    sites = []
    for _ in range(3):
        a = Adapter(names=["adapter_0", "adapter_1", "adapter_2"], device="cuda", rank=r)
        with torch.no_grad():
            for p in a.parameters():
                p.copy_(torch.randn(p.shape, device="cuda", generator=g) * (0.1 if p.dim() == 1 else 0.05))
        if gating:
            a.activate_gating(); a.set_active_adapter("adapter_0")
        else:
            a.deactivate_gating(); a.set_active_adapter("adapter_1")
        sites.append(a)
    x0 = torch.randn(4, 100, 768, device="cuda", generator=g).to(torch.bfloat16)
    r0 = torch.randn(4, 100, 768, device="cuda", generator=g).to(torch.bfloat16)
    dy = torch.randn(4, 100, 768, device="cuda", generator=g).to(torch.bfloat16)

    def run(defer):
        for a in sites:
            a.zero_grad(set_to_none=True)
        x, res = x0.clone().requires_grad_(True), r0.clone().requires_grad_(True)
        h = x
        for a in sites:
            h = a(h, res)
        n0 = ops.launch_count
        if defer:
            with ops.deferred_wgrad() as q:
                h.backward(dy)
                queued = len(q.groups)
        else:
            h.backward(dy)
            queued = 0
        torch.cuda.synchronize()
        grads = {f"{i}.{n}": (None if p.grad is None else p.grad.clone()) for i, a in enumerate(sites) for n, p in a.named_parameters()}
        return x.grad.clone(), res.grad.clone(), grads, queued, ops.launch_count - n0

    dx_r, dr_r, ref, _, n_ref = run(False)
    dx_d, dr_d, got, queued, n_def = run(True)
    assert torch.equal(dx_r, dx_d) and torch.equal(dr_r, dr_d)
    assert queued == (0 if r == 512 else 3 * max(1, r // 128))
    assert n_def < n_ref or r == 512
    trained = "adapter_0" if gating else "adapter_1"
    for n in ref:
        if trained in n:
            assert got[n] is not None and got[n].is_contiguous() and got[n].shape == ref[n].shape
            assert relerr(got[n].cpu().numpy(), ref[n].cpu().numpy()) < 1e-5, n
        else:
            assert ref[n] is None and got[n] is None, n
