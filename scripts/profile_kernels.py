"""The DAT kernels alone at the step's shapes and at the 12-site batched size, inside a profiler
range, for `ncu --set full -k regex:dat_ --profile-from-start off ...`."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from feddat_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
g = torch.Generator(device=dev).manual_seed(0)
D, r = 768, int(sys.argv[1]) if len(sys.argv) > 1 else 128
Ms = [int(a) for a in sys.argv[2:]] or [5920, 71040]


def mk(nb):
    return ops.pack_weights([[torch.randn(r, D, device=dev, generator=g) * 0.02, torch.zeros(r, device=dev),
                              torch.randn(D, r, device=dev, generator=g) * 0.02, torch.zeros(D, device=dev)]
                             for _ in range(nb)])


pk2, pk1 = mk(2), mk(1)
for M in Ms:
    x = torch.randn(M, D, device=dev, generator=g).to(torch.bfloat16)
    dy = torch.randn(M, D, device=dev, generator=g).to(torch.bfloat16)
    # the product path for ReLU: the forward saves the hidden, the backward does not recompute it
    for _ in range(2):
        _, h2 = ops.dat_forward(x, x, pk2, 0.5, save_hidden=True)
        _, h1 = ops.dat_forward(x, x, pk1, 1.0, save_hidden=True)
        ops.dat_backward(x, dy, pk2, 0.5, train_slice=(0, r), hidden=h2)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    ops.dat_forward(x, x, pk2, 0.5, save_hidden=True)
    ops.dat_forward(x, x, pk1, 1.0, save_hidden=True)
    ops.dat_backward(x, dy, pk2, 0.5, train_slice=(0, r), hidden=h2)
    ops.dat_backward(x, dy, pk1, 1.0, train_slice=(0, r), hidden=h1)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
