"""Fused first-GEMM + exact-GELU kernels of the frozen MLP vs the unfused pair (cuBLAS + streaming GELU kernel),
at the benchmarked step's shape.  CUDA events, cold inputs (rotation > 2x L2)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from feddat_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
M, N, K = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (11840, 3072, 768)
w = (torch.randn(N, K, device=dev, generator=g) * 0.05).to(torch.bfloat16)
b = (torch.randn(N, device=dev, generator=g) * 0.1).to(torch.bfloat16)
w2 = (torch.randn(K, N, device=dev, generator=g) * 0.05).to(torch.bfloat16)      # ViltOutput.dense.weight layout
w2t = w2.t().contiguous()
sets = []
for _ in range(4):
    a = torch.randn(M, K, device=dev, generator=g).to(torch.bfloat16)
    dy = torch.randn(M, K, device=dev, generator=g).to(torch.bfloat16)
    pre = torch.addmm(b, a, w.t())
    sets.append((a, dy, pre))


def timeit(fn, iters=10):
    for i in range(3):
        fn(*sets[i % 4])
    ts = []
    for i in range(iters):
        torch.cuda._sleep(1_000_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(*sets[(3 + i) % 4]); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sum(ts) / len(ts)


flops = 2 * M * N * K
b32 = b.float()
t = timeit(lambda a, dy, pre: ops.mlp_fc1_gelu(a, w, b32))
print(f"fused fwd  (GEMM + bias + GELU, writes pre and act): {t:7.1f} us  {flops / t / 1e6:7.1f} TFLOP/s")
t = timeit(lambda a, dy, pre: ops.gelu_fwd(torch.addmm(b, a, w.t())))
print(f"unfused fwd (cuBLAS addmm + gelu kernel):            {t:7.1f} us")
t = timeit(lambda a, dy, pre: torch.addmm(b, a, w.t()))
print(f"  cuBLAS addmm alone:                                {t:7.1f} us  {flops / t / 1e6:7.1f} TFLOP/s")
t = timeit(lambda a, dy, pre: ops.mlp_fc2_dgelu(dy, w2t, pre))
print(f"fused bwd  (GEMM * gelu'(pre)):                      {t:7.1f} us  {flops / t / 1e6:7.1f} TFLOP/s")
t = timeit(lambda a, dy, pre: ops.gelu_bwd(torch.mm(dy, w2), pre))
print(f"unfused bwd (cuBLAS mm + gelu_bwd kernel):           {t:7.1f} us")
