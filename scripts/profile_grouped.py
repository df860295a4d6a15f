"""The DAT launches of ONE adapter site of the batched MKD schedule at BASELINE configs[1] sizes, for ncu:
grouped forward, grouped backward data gradient, grouped weight gradient (5 920 gating rows with R = 256 +
5 920 adapter_1 rows with R = 128), plus the batched weight pack of 12 sites.
    ncu --set full --clock-control none --import-source on -o gpurun_out/r2_grouped python scripts/profile_grouped.py"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from feddat_b200 import ops  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 5920
r = 128
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)


def branches(nb):
    return [[torch.randn(r, 768, device=dev, generator=g) * 0.02, torch.zeros(r, device=dev),
             torch.randn(768, r, device=dev, generator=g) * 0.02, torch.zeros(768, device=dev)] for _ in range(nb)]


specs = []
for _ in range(12):
    b2, b1 = branches(2), branches(1)
    specs.append(ops.PackSpec([b[0] for b in b2], [b[1] for b in b2], [b[2] for b in b2], [b[3] for b in b2]))
    specs.append(ops.PackSpec([b[0] for b in b1], [b[1] for b in b1], [b[2] for b in b1], [b[3] for b in b1]))
packs = ops.pack_weights_batched(specs)
pk2, pk1 = packs[0], packs[1]
x = torch.randn(2 * M, 768, device=dev, generator=g).to(torch.bfloat16)
dy = torch.randn(2 * M, 768, device=dev, generator=g).to(torch.bfloat16)
y, dx = torch.empty_like(x), torch.empty_like(x)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
for it in range(3):
    flush.zero_()            # evict: the kernels then read their inputs from HBM like inside a train step
    (_, h2), (_, h1) = ops.dat_forward_grouped([dict(x=x[:M], res=x[:M], w=pk2, scale=0.5, out=y[:M], save_hidden=True),
                                                dict(x=x[M:], res=x[M:], w=pk1, scale=1.0, out=y[M:], save_hidden=True)])
    flush.zero_()
    ops.dat_backward_grouped([dict(x=x[:M], dy=dy[:M], w=pk2, scale=0.5, train_slice=(0, r), hidden=h2, dx_out=dx[:M]),
                              dict(x=x[M:], dy=dy[M:], w=pk1, scale=1.0, train_slice=(0, r), hidden=h1, dx_out=dx[M:])])
torch.cuda.synchronize()
print("ok")
