"""The DAT launches of the batched MKD schedule at BASELINE configs[1] sizes, for ncu: ONE adapter site's grouped forward
and grouped backward data gradient (5 920 gating rows with R = 256 + 5 920 adapter_1 rows with R = 128, inputs evicted
from L2 first), and the deferred weight-gradient launch over all 12 sites (24 groups x 6 chunks = 144 CTAs).
    ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:dat_ \
        -o gpurun_out/r2_grouped_v2 python scripts/profile_grouped.py"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from feddat_b200 import ops  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 5920
r, SITES = 128, 12
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)


def mk(nb):
    return ops.pack_weights([[torch.randn(r, 768, device=dev, generator=g) * 0.02, torch.zeros(r, device=dev),
                              torch.randn(768, r, device=dev, generator=g) * 0.02, torch.zeros(768, device=dev)]
                             for _ in range(nb)])


pk2, pk1 = mk(2), mk(1)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)


def forward(x, y):
    return ops.dat_forward_grouped([dict(x=x[:M], res=x[:M], w=pk2, scale=0.5, out=y[:M], save_hidden=True),
                                    dict(x=x[M:], res=x[M:], w=pk1, scale=1.0, out=y[M:], save_hidden=True)])


def backward(x, dy, h2, h1, dx):
    return ops.dat_backward_grouped([dict(x=x[:M], dy=dy[:M], w=pk2, scale=0.5, train_slice=(0, r), hidden=h2, dx_out=dx[:M]),
                                     dict(x=x[M:], dy=dy[M:], w=pk1, scale=1.0, train_slice=(0, r), hidden=h1, dx_out=dx[M:])],
                                    allow_defer=True)


sites = []
for _ in range(SITES):
    x = torch.randn(2 * M, 768, device=dev, generator=g).to(torch.bfloat16)
    dy = torch.randn(2 * M, 768, device=dev, generator=g).to(torch.bfloat16)
    y = torch.empty_like(x)
    (_, h2), (_, h1) = forward(x, y)
    sites.append((x, dy, h2, h1, y))
for it in range(3):
    last = it == 2
    with ops.deferred_wgrad() as q:
        for s in sites[:-1]:
            backward(*s)
        x, dy, h2, h1, y = sites[-1]
        flush.zero_()            # evict: the kernels then read their inputs from HBM like inside a train step
        torch.cuda.synchronize()
        if last:
            torch.cuda.profiler.start()
        forward(x, y)
        torch.cuda.synchronize()
        if last:
            torch.cuda.profiler.stop()
        flush.zero_()
        torch.cuda.synchronize()
        if last:
            torch.cuda.profiler.start()
        backward(*sites[-1])
        q.flush()
        torch.cuda.synchronize()
        if last:
            torch.cuda.profiler.stop()
print("ok")
