"""Small driver for compute-sanitizer (memcheck / racecheck / synccheck) over the DAT kernels:
    compute-sanitizer --tool memcheck  python scripts/sanitize_kernels.py 1 129 5920 25003
    compute-sanitizer --tool racecheck python scripts/sanitize_kernels.py 1 129 5920
Runs, per row count M: single-group forward / saved-mode backward (dat_fused_kernel or dat_pipe_kernel by
size), the grouped launches over [M gating rows | M adapter_1 rows], the deterministic weight-gradient kernel,
the batched weight pack and both MKD heads; checks the results against torch so that a sanitizer run is also a
functional run."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from feddat_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
r = 64


def branches(nb):
    return [[torch.randn(r, 768, device=dev, generator=g) * 0.05, torch.randn(r, device=dev, generator=g) * 0.1,
             torch.randn(768, r, device=dev, generator=g) * 0.05, torch.randn(768, device=dev, generator=g) * 0.1]
            for _ in range(nb)]


def ref_fwd(x, brs, scale):
    up = 0
    for dw, db, uw, ub in brs:
        h = torch.relu(x.float() @ dw.to(torch.bfloat16).float().T + db).to(torch.bfloat16).float()
        up = up + h @ uw.to(torch.bfloat16).float().T + ub
    return x.float() + (scale * up).to(torch.bfloat16).float()


for M in [int(a) for a in sys.argv[1:]] or [129]:
    b2, b1 = branches(2), branches(1)
    pk2, pk1 = ops.pack_weights_batched([ops.PackSpec(*[[b[i] for b in bs] for i in range(4)]) for bs in (b2, b1)])
    x = torch.randn(2 * M, 768, device=dev, generator=g).to(torch.bfloat16)
    dy = torch.randn(2 * M, 768, device=dev, generator=g).to(torch.bfloat16)
    y, dx = torch.empty_like(x), torch.empty_like(x)
    (_, h2), (_, h1) = ops.dat_forward_grouped([dict(x=x[:M], res=x[:M], w=pk2, scale=0.5, out=y[:M], save_hidden=True),
                                                dict(x=x[M:], res=x[M:], w=pk1, scale=1.0, out=y[M:], save_hidden=True)])
    res = ops.dat_backward_grouped([dict(x=x[:M], dy=dy[:M], w=pk2, scale=0.5, train_slice=(0, r), hidden=h2, dx_out=dx[:M]),
                                    dict(x=x[M:], dy=dy[M:], w=pk1, scale=1.0, train_slice=(0, r), hidden=h1, dx_out=dx[M:])])
    y1, hh = ops.dat_forward(x[:M], x[:M], pk2, 0.5, save_hidden=True)
    dx1, gr1 = ops.dat_backward(x[:M], dy[:M], pk2, 0.5, train_slice=(0, r), hidden=hh)
    torch.cuda.synchronize()
    e_y = ((y[:M].float() - ref_fwd(x[:M], b2, 0.5)).abs().max() / y[:M].float().abs().max()).item()
    e_y1 = ((y[M:].float() - ref_fwd(x[M:], b1, 1.0)).abs().max() / y[M:].float().abs().max()).item()
    assert e_y < 1e-2 and e_y1 < 1e-2, (e_y, e_y1)
    assert torch.equal(y1, y[:M]) and torch.equal(dx1, dx[:M])
    print(f"M={M}: grouped / single forward + backward ok (fwd err {e_y:.1e} / {e_y1:.1e})")

# ---- deferred weight gradients: 13 sites x 2 row groups = 26 groups -> one 24-group launch (no row splits) + one
# 2-group launch (row splits, two-stage reduction), against the per-site launches
Md = 333
sites = []
for _ in range(13):
    b2, b1 = branches(2), branches(1)
    pk2, pk1 = ops.pack_weights_batched([ops.PackSpec(*[[b[i] for b in bs] for i in range(4)]) for bs in (b2, b1)])
    x = torch.randn(2 * Md, 768, device=dev, generator=g).to(torch.bfloat16)
    dy = torch.randn(2 * Md, 768, device=dev, generator=g).to(torch.bfloat16)
    (_, h2), (_, h1) = ops.dat_forward_grouped([dict(x=x[:Md], res=x[:Md], w=pk2, scale=0.5, save_hidden=True),
                                                dict(x=x[Md:], res=x[Md:], w=pk1, scale=1.0, save_hidden=True)])
    sites.append((x, dy, pk2, pk1, h2, h1))


def site_bwd(s_):
    x, dy, pk2, pk1, h2, h1 = s_
    return ops.dat_backward_grouped([dict(x=x[:Md], dy=dy[:Md], w=pk2, scale=0.5, train_slice=(0, r), hidden=h2),
                                     dict(x=x[Md:], dy=dy[Md:], w=pk1, scale=1.0, train_slice=(0, r), hidden=h1)],
                                    allow_defer=True)


with ops.deferred_wgrad():
    deferred = [site_bwd(s_) for s_ in sites]
direct = [site_bwd(s_) for s_ in sites]
torch.cuda.synchronize()
for a_, b_ in zip(deferred, direct):
    for (_, ga), (_, gb) in zip(a_, b_):
        for ta, tb in zip(ga, gb):
            assert ((ta - tb).abs().max() / tb.abs().max().clamp_min(1e-20)).item() < 1e-5
print("deferred weight gradients (26 groups): ok")

lg = torch.randn(32, 100, device=dev, generator=g)
ops.mkd_loss(lg, torch.randn(32, 100, device=dev, generator=g), (torch.rand(32, 100, device=dev, generator=g) < 0.02).float(), 2.0)
sc = torch.randn(5, 4, 3202, device=dev, generator=g).to(torch.bfloat16)
lab = torch.randint(0, 3202, (5, 4), device=dev, generator=g)
lab[1, 2:] = -100
ops.mkd_ce_loss(sc, torch.randn(5, 4, 3202, device=dev, generator=g).to(torch.bfloat16)[:, :-1], lab,
                torch.rand(5, device=dev, generator=g), 2.0)
torch.cuda.synchronize()
print("mkd heads ok")


# ---- the frozen block's own kernels: GEMM + GELU epilogues, short-sequence attention (forward, statistics, backward)
for M, N, K in ((257, 512, 128), (5920, 3072, 768)):
    a = torch.randn(M, K, device=dev, generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev, generator=g) * 0.05).to(torch.bfloat16)
    b = torch.randn(N, device=dev, generator=g) * 0.1
    pre, act = ops.mlp_fc1_gelu(a, w, b)
    dpre = ops.mlp_fc2_dgelu(a, w, pre)
    torch.cuda.synchronize()
    ref = torch.nn.functional.gelu(torch.addmm(b, a.float(), w.float().t()))
    e = ((act.float() - ref).abs().max() / ref.abs().max()).item()
    assert e < 1e-2, e
    print(f"mlp gemm M={M} N={N} K={K}: ok (act err {e:.1e})")
for B, S, H in ((2, 185, 3), (8, 185, 12), (3, 64, 2), (2, 129, 1)):
    q, k, v, do = (torch.randn(B, S, H, 64, device=dev, generator=g).to(torch.bfloat16) for _ in range(4))
    o, lse = ops.attn_fwd(q, k, v, 0.125)
    dq, dk, dv = ops.attn_bwd(do, q, k, v, o, lse, 0.125)
    torch.cuda.synchronize()
    qf, kf, vf = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))
    ref = (torch.softmax(qf @ kf.transpose(-1, -2) * 0.125, -1) @ vf).permute(0, 2, 1, 3)
    e = ((o.float() - ref).abs().max() / ref.abs().max()).item()
    assert e < 1e-2 and torch.isfinite(dq.float()).all() and torch.isfinite(dk.float()).all() and torch.isfinite(dv.float()).all(), e
    print(f"attention B={B} S={S} H={H}: forward + backward ok (out err {e:.1e})")
