"""BASELINE configs[4]: adapter-rank sweep r in {32, 48, 64, 128, 256, 512} on one B200 through the
drop-in ``Adapter`` module (gating mode: adapter_0 trainable + adapter_2 frozen, residual = input): fused
DAT forward + backward time, achieved algorithmic TFLOP/s / GB/s and fraction of the roofline
(20 d r FLOP and 10 d bytes per row, fwd + bwd gating)."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from feddat_b200.modeling.adapter import Adapter  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
peaks = bench.read_peaks()
D = 768
rows = []
for M in (5920, 71040):
    n_sets = max(2, -(-300_000_000 // (2 * M * D * 2)))
    g = torch.Generator(device=dev).manual_seed(M)
    sets = [(torch.randn(M, D, device=dev, generator=g).to(torch.bfloat16),
             torch.randn(M, D, device=dev, generator=g).to(torch.bfloat16)) for _ in range(n_sets)]
    for r in (32, 48, 64, 128, 256, 512):
        ad = Adapter(names=["adapter_0", "adapter_1", "adapter_2"], device="cuda", rank=r)
        ad.activate_gating(); ad.set_active_adapter("adapter_0")

        def step(x, dy):
            xin = x.detach().requires_grad_(True)
            y = ad(xin, xin)
            y.backward(dy)

        for i in range(3):
            step(*sets[i % n_sets])
        ts = []
        for i in range(10):
            x, dy = sets[(3 + i) % n_sets]
            torch.cuda._sleep(3_000_000)       # ~1.5 ms: the whole module call (pack, fwd, autograd, dgrad, wgrad) is enqueued behind it
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); step(x, dy); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        t = sum(ts) / len(ts)
        flops, nbytes = 20 * D * r * M, 10 * D * M
        t_roof = max(flops / (peaks["tf_burst"] * 1e12), nbytes / (peaks["hbm_gbs"] * 1e9))
        rows.append({"M": M, "r": r, "us": round(t * 1e6, 1), "tflops": round(flops / t / 1e12, 1),
                     "gbs": round(nbytes / t / 1e9, 1), "bound": "tensor" if flops / (peaks["tf_burst"] * 1e12) >= nbytes / (peaks["hbm_gbs"] * 1e9) else "hbm",
                     "frac": round(t_roof / t, 3)})
        print(rows[-1], flush=True)
print(json.dumps(rows))
