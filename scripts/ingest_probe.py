"""Per-SM TMA ingest sweep (csrc/probe.cu probe_ingest_kernel): ring depth x box height x producers x grid."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from feddat_b200 import _lib  # noqa: E402

lib = _lib.load_debug()
dev = torch.device("cuda", 0)
n_rows = 32768
clk = torch.zeros(1, dtype=torch.int64, device=dev)
for stride in (768,):
    buf = torch.randn(n_rows, stride, device=dev).to(torch.bfloat16)      # 4 MB / 50 MB: L2 resident
    for grid in (8, 48, 148):
        for box_rows in (64, 128, 256):
            for ns in (2, 4, 8, 12):
                if grid == 48 and box_rows != 128:
                    continue
                if ns * box_rows * 128 > 200 * 1024:
                    continue
                for n_prod in (1, 2):
                    iters = 4000 * 128 // box_rows
                    args = (_lib.ptr(buf), n_rows, stride, iters, grid, ns, box_rows, n_prod, _lib.ptr(clk), _lib.stream_ptr())
                    _lib.check(lib.feddat_probe_ingest(*args))
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    _lib.check(lib.feddat_probe_ingest(*args))
                    e1.record()
                    torch.cuda.synchronize()
                    sec = e0.elapsed_time(e1) * 1e-3
                    per_sm = iters * box_rows * 128 / sec / 1e9
                    print(f"stride={stride:4d} grid={grid:3d} box={box_rows:3d}x64 ns={ns:2d} prod={n_prod}: "
                          f"{per_sm:7.1f} GB/s per SM  {per_sm * grid / 1e3:6.2f} TB/s chip   issue {clk.item()} clk/load")
