"""Weight gradients of all 12 adapter sites of the batched MKD schedule (BASELINE configs[1]: per site 5 920 gating
rows with R = 256 + 5 920 adapter_1 rows, r_t = 128 each): ONE deferred launch (ops.deferred_wgrad: 24 groups x 6
chunks = 144 CTAs, no row splits) against the twelve per-site launches it replaces, CUDA events, inputs cold (the
twelve sites' X / dY / H / dP are 0.5 GB: nothing survives in the 126 MB L2 from one launch to the next).
    python scripts/bench_deferred_wgrad.py            # timings
    ncu --set full --clock-control none -k regex:dat_wgrad -o gpurun_out/r2_wgrad12 python scripts/bench_deferred_wgrad.py ncu"""
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from feddat_b200 import ops  # noqa: E402

NCU = len(sys.argv) > 1 and sys.argv[1] == "ncu"
M, r, SITES = 5920, 128, 12
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)


def mk(nb):
    return ops.pack_weights([[torch.randn(r, 768, device=dev, generator=g) * 0.02, torch.zeros(r, device=dev),
                              torch.randn(768, r, device=dev, generator=g) * 0.02, torch.zeros(768, device=dev)]
                             for _ in range(nb)])


pk2, pk1 = mk(2), mk(1)
sites = []
for _ in range(SITES):
    x = torch.randn(2 * M, 768, device=dev, generator=g).to(torch.bfloat16)
    dy = torch.randn(2 * M, 768, device=dev, generator=g).to(torch.bfloat16)
    y = torch.empty_like(x)
    (_, h2), (_, h1) = ops.dat_forward_grouped([dict(x=x[:M], res=x[:M], w=pk2, scale=0.5, out=y[:M], save_hidden=True),
                                                dict(x=x[M:], res=x[M:], w=pk1, scale=1.0, out=y[M:], save_hidden=True)])
    sites.append((x, dy, h2, h1, y))


def backward(site):
    x, dy, h2, h1, dx = site
    return ops.dat_backward_grouped([dict(x=x[:M], dy=dy[:M], w=pk2, scale=0.5, train_slice=(0, r), hidden=h2, dx_out=dx[:M]),
                                     dict(x=x[M:], dy=dy[M:], w=pk1, scale=1.0, train_slice=(0, r), hidden=h1, dx_out=dx[M:])],
                                    allow_defer=True)


def timed(fn):
    torch.cuda._sleep(2_000_000)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3


res = {"deferred_flush": [], "deferred_total": [], "per_site_total": []}
for it in range(3 if NCU else 8):
    with ops.deferred_wgrad() as q:
        t_d = timed(lambda: [backward(s) for s in sites])          # 12 data-gradient launches
        t_f = timed(q.flush)                                        # ONE weight-gradient launch
    t_p = timed(lambda: [backward(s) for s in sites])               # 12 x (dgrad + wgrad)
    if it >= 2:
        res["deferred_flush"].append(t_f)
        res["deferred_total"].append(t_d + t_f)
        res["per_site_total"].append(t_p)
if not NCU:
    out = {k: round(statistics.mean(v), 1) for k, v in res.items()}
    alg = SITES * (2 * 2 * M * 768 * 2 + 2 * 2 * M * r * 2)        # X, dY + H_t, dP_t of both groups
    out["wgrad_algorithmic_MB"] = round(alg / 1e6, 1)
    out["deferred_flush_GBs"] = round(alg / out["deferred_flush"] / 1e3, 1)
    out["per_site_wgrad_us_now"] = round(out["deferred_flush"] / SITES, 2)
    print(out)
