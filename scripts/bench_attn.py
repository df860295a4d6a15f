"""Short-sequence attention kernel vs F.scaled_dot_product_attention (cuDNN / flash backends) at the benchmarked
step's shape.  CUDA events, rotating input sets."""
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from feddat_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
B, S, H, D = (int(a) for a in sys.argv[1:5]) if len(sys.argv) > 4 else (64, 185, 12, 64)
sets = []
for _ in range(6):
    q, k, v, do = ((torch.randn(B * S, H * D, device=dev, generator=g)).to(torch.bfloat16).view(B, S, H, D) for _ in range(4))
    sets.append((q, k, v, do))


def timeit(fn, iters=12):
    for i in range(3):
        fn(*sets[i % 6])
    ts = []
    for i in range(iters):
        torch.cuda._sleep(1_000_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(*sets[(3 + i) % 6]); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sum(ts) / len(ts)


t = timeit(lambda q, k, v, do: ops.attn_fwd(q, k, v, 0.125))
print(f"own fwd:                 {t:7.1f} us")
t = timeit(lambda q, k, v, do: F.scaled_dot_product_attention(q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3)))
print(f"torch SDPA fwd:          {t:7.1f} us")
if hasattr(ops, "attn_bwd"):
    o, lse = ops.attn_fwd(*sets[0][:3], 0.125)
    t = timeit(lambda q, k, v, do: ops.attn_bwd(do, q, k, v, o, lse, 0.125))
    print(f"own bwd:                 {t:7.1f} us")


def torch_fb(q, k, v, do):
    q, k, v = (t.detach().permute(0, 2, 1, 3).requires_grad_(True) for t in (q, k, v))
    out = F.scaled_dot_product_attention(q, k, v)
    out.backward(do.permute(0, 2, 1, 3))


t = timeit(torch_fb)
print(f"torch SDPA fwd + bwd:    {t:7.1f} us")
