"""One ALBEF train step (BASELINE configs[2]) inside a cudaProfilerStart/Stop range, for
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ... python scripts/profile_albef_step.py"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from feddat_b200.synthetic import albef_to_device, make_albef_batch  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
tr, wrapped, opt, sched = bench.build_albef_client(0, dev)
batches = [albef_to_device(make_albef_batch(16, 384, seed=i, client=0), dev) for i in range(3)]
for i in range(2):
    tr.train_step(wrapped, i, batches[i], opt, sched)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.train_step(wrapped, 2, batches[2], opt, sched)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one ALBEF train step")
