"""L2 -> SM TMA streaming bandwidth, unicast vs cluster multicast (see csrc/probe.cu)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from feddat_b200 import _lib  # noqa: E402

lib = _lib.load_debug()
dev = torch.device("cuda", 0)
for n_boxes in (48, 4096):          # 768 KB (weights-like, L2 resident) and 64 MB
    buf = torch.randn(n_boxes * 128, 64, device=dev).to(torch.bfloat16)
    for cluster in (1, 2, 4, 8):
        grid = 148 - 148 % cluster
        iters = 2000
        for _ in range(2):
            _lib.check(lib.feddat_probe_l2bw(_lib.ptr(buf), n_boxes, iters, grid, cluster, _lib.stream_ptr()))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.feddat_probe_l2bw(_lib.ptr(buf), n_boxes, iters, grid, cluster, _lib.stream_ptr()))
        e1.record()
        torch.cuda.synchronize()
        sec = e0.elapsed_time(e1) * 1e-3
        smem_bytes = grid * iters * 16384
        print(f"boxes={n_boxes:5d} cluster={cluster} grid={grid}: {sec * 1e6:8.1f} us  smem fill {smem_bytes / sec / 1e12:6.2f} TB/s"
              f"  ({smem_bytes / sec / grid / 1e9:6.1f} GB/s per SM)   L2 reads {smem_bytes / cluster / sec / 1e12:6.2f} TB/s")
