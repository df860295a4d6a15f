"""LayerNorm / GELU kernel timings at the step's shapes (cold inputs, CUDA events) next to torch's."""
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from feddat_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
M = 5920
n_sets = 12
g = torch.Generator(device=dev).manual_seed(0)
big = [torch.randn(M, 3072, device=dev, generator=g).to(torch.bfloat16) for _ in range(n_sets)]
dyb = [torch.randn(M, 3072, device=dev, generator=g).to(torch.bfloat16) for _ in range(n_sets)]
x = [torch.randn(M, 768, device=dev, generator=g).to(torch.bfloat16) for _ in range(4 * n_sets)]
w = torch.ones(768, device=dev, dtype=torch.bfloat16)
b = torch.zeros(768, device=dev, dtype=torch.bfloat16)


def timeit(fn, n):
    for i in range(3):
        fn(i)
    ts = []
    for i in range(3, 3 + n):
        torch.cuda._sleep(100_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(i); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sum(ts) / len(ts)


print("gelu fwd   ours %.1f us   torch %.1f us   (72.7 MB)" % (
    timeit(lambda i: ops.gelu_fwd(big[i % n_sets]), 9), timeit(lambda i: F.gelu(big[i % n_sets]), 9)))
print("gelu bwd   ours %.1f us   (109 MB)" % timeit(lambda i: ops.gelu_bwd(dyb[i % n_sets], big[i % n_sets]), 9))
print("add+LN fwd ours %.1f us   torch add + layer_norm %.1f us   (45 MB incl. the pre-biased stream)" % (
    timeit(lambda i: ops.layer_norm_fwd(x[i % 48], x[(i + 7) % 48], w, b, 1e-12, bias2=b), 20),
    timeit(lambda i: F.layer_norm(x[i % 48] + x[(i + 7) % 48], (768,), w, b, 1e-12), 20)))
y, s, mean, rstd, _ = ops.layer_norm_fwd(x[0], x[1], w, b, 1e-12)
print("LN bwd     ours %.1f us   (36 MB)" % timeit(lambda i: ops.layer_norm_bwd(x[i % 48], x[(i + 5) % 48], s, w, mean, rstd), 20))
