"""Timeline of CTA 0's first items inside attn_fwd_kernel (debug twin: FEDDAT_DEBUG_LIB=1 python scripts/trace_attn.py)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from feddat_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
B, S, H, D = 64, 185, 12, 64
q, k, v = (torch.randn(B * S, H * D, device=dev, generator=g).to(torch.bfloat16).view(B, S, H, D) for _ in range(3))
lib = _lib.load_debug()
for _ in range(3):
    ops.attn_fwd(q, k, v, 0.125)
buf = torch.zeros(256, dtype=torch.int64, device=dev)
lib.feddat_debug_set_trace(_lib.ptr(buf))
ops.attn_fwd(q, k, v, 0.125)
torch.cuda.synchronize()
lib.feddat_debug_set_trace(None)
t = buf.cpu().view(16, 16).tolist()
names = ["ctl: Q,K landed", "ctl: S done", "ctl: V landed", "ctl: P ready", "ctl: O done", "smx: S seen", "smx: row loaded + max",
         "smx: max exchanged", "smx: P stored, arrived", "epi: start", "epi: sums exchanged", "epi: O seen", "smx: epilogue(i-1) done"]
t0 = min(x for row in t for x in row if x)
for i in range(12):
    print(f"item {i}: " + "  ".join(f"{names[e].split(':')[0]}{e}={(t[i][e] - t0) / 1e3:6.2f}" for e in range(13) if t[i][e]))
print("events:", {e: n for e, n in enumerate(names)})

# ---- backward
do = torch.randn(B, S, H, D, device=dev, generator=g).to(torch.bfloat16)
o, lse = ops.attn_fwd(q, k, v, 0.125)
for _ in range(3):
    ops.attn_bwd(do, q, k, v, o, lse, 0.125)
buf.zero_()
lib.feddat_debug_set_trace(_lib.ptr(buf))
ops.attn_bwd(do, q, k, v, o, lse, 0.125)
torch.cuda.synchronize()
lib.feddat_debug_set_trace(None)
t = buf.cpu().view(16, 16).tolist()
names = ["mma: S^T issued", "mma: dP^T issued", "mma: dS^T seen", "mma: dV dK dQ issued", "", "smx: S^T seen", "smx: P^T packed", "smx: dP^T seen",
         "smx: dS^T out", "epi: dV dK done seen", "epi: dK stored, Y free", "epi: dQ stored"]
t0 = min(x for row in t for x in row if x)
print("backward, key tile steps g (two per item):")
for i in range(12):
    print(f"g {i}: " + "  ".join(f"{names[e].split(':')[0]}{e}={(t[i][e] - t0) / 1e3:6.2f}" for e in range(12) if t[i][e]))
print("events:", {e: n for e, n in enumerate(names) if n})
