"""A/B of module-level switches on the graphed train step (BASELINE configs[1]), interleaved in ONE process so that
clocks / power state are shared:   python scripts/ab_step.py task_trainer.DEFER_WGRAD [vilt.FUSE_ATTENTION ...]
For every named flag: capture one graph with the flag True and one with it False (everything else at its default),
then alternate blocks of 40 replays and report the mean ms / step of each."""
import importlib
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from feddat_b200.synthetic import make_vilt_batch, to_device  # noqa: E402
from feddat_b200.train.graphed import GraphedTrainStep  # noqa: E402

MODS = {"task_trainer": "feddat_b200.train.task_trainer", "vilt": "feddat_b200.modeling.vilt",
        "fused_ln": "feddat_b200.modeling.fused_ln", "ops": "feddat_b200.ops"}
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
batches = [to_device(make_vilt_batch(bench.B, bench.T, bench.H, bench.C, seed=i, client=0), dev) for i in range(4)]


def graphed(flag, value):
    mod, name = flag.split(".")
    m = importlib.import_module(MODS[mod])
    old = getattr(m, name)
    setattr(m, name, value)
    c = bench.build_client(0, dev)
    g = GraphedTrainStep(c.trainer, c.wrapped, c.opt, c.sched, batches[0], warmup=2)
    for i in range(4):
        g(batches[i % 4])
    torch.cuda.synchronize()
    setattr(m, name, old)
    return g


def block(g, n=40):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        g(batches[i % 4])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for flag in sys.argv[1:]:
    on, off = graphed(flag, True), graphed(flag, False)
    t_on, t_off = [], []
    for _ in range(5):
        t_on.append(block(on))
        t_off.append(block(off))
    print(f"{flag}: True {statistics.mean(t_on):.3f} ms  False {statistics.mean(t_off):.3f} ms  "
          f"(blocks on {[round(t, 3) for t in t_on]} off {[round(t, 3) for t in t_off]})", flush=True)
    del on, off
