"""DAT kernel timings alone (bench.kernel_rooflines): CUDA events, L2 flush between launches."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

torch.cuda.set_device(0)
out = bench.kernel_rooflines(bench.read_peaks(), torch.device("cuda", 0))
for k, v in out.items():
    print(f"{k:22s} {v['us']:8.2f} us  {v['tflops']:7.1f} TFLOP/s  {v['gbs']:7.1f} GB/s  {100 * v['frac_of_roofline']:5.1f}% of {v['bound']} roofline")
print(json.dumps(out))
