"""One train step inside a cudaProfilerStart/Stop range, for
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ...
(launch list of the step) -- see profiles/README.md."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from feddat_b200.synthetic import make_vilt_batch, to_device  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
c = bench.build_client(0, dev)
batches = [to_device(make_vilt_batch(bench.B, bench.T, bench.H, bench.C, seed=i, client=0), dev) for i in range(4)]
for i in range(3):
    c.trainer.train_step(c.wrapped, i, batches[i], c.opt, c.sched)
torch.cuda.synchronize()
torch.cuda.profiler.start()
c.trainer.train_step(c.wrapped, 3, batches[3], c.opt, c.sched)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one train step")
