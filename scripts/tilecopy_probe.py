"""HBM bandwidth of the DAT kernels' access pattern without compute (csrc/probe.cu probe_tilecopy_kernel),
next to torch's linear copy."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from feddat_b200 import _lib  # noqa: E402

lib = _lib.load_debug()
dev = torch.device("cuda", 0)
M = 71040 * 4            # 436 MB per tensor: well beyond L2
src = torch.randn(M, 768, device=dev).to(torch.bfloat16)
dst = torch.empty_like(src)
nbytes = src.numel() * 2


def timed(fn, n=5):
    fn()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return min(ts)


t = timed(lambda: dst.copy_(src))
print(f"torch copy_            : {2 * nbytes / t / 1e12:5.2f} TB/s (read + write)")
for ns in (4, 8, 12):
    for grid in (148, 296):
        g = min(grid, 148) if ns > 6 else grid      # two CTAs per SM only fit with small rings
        for ro in (1, 0):
            t = timed(lambda: _lib.check(lib.feddat_probe_tilecopy(_lib.ptr(src), _lib.ptr(dst), M, g, ns, ro, _lib.stream_ptr())))
            moved = nbytes * (1 if ro else 2)
            print(f"tile copy ns={ns:2d} grid={g:3d} {'read only ' if ro else 'read+write'}: {moved / t / 1e12:5.2f} TB/s")
