"""Debug aid: eager vs graphed train_step losses, step by step."""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from test_train_step_gpu import build_trainer  # noqa: E402
from feddat_b200.synthetic import make_vilt_batch, to_device  # noqa: E402
from feddat_b200.train.graphed import GraphedTrainStep  # noqa: E402
from feddat_b200.train.prepare import default_args, place_on_gpu, prepare_model  # noqa: E402
from feddat_b200.train.task_trainer import get_polynomial_decay_schedule_with_warmup  # noqa: E402


def build():
    torch.manual_seed(7)
    model = prepare_model(default_args(ordered_cl_tasks=["art"], adapter_rank=32), place=False)
    place_on_gpu(model)
    tr = build_trainer(model, 1e-3, 20, "art", temp=2.0)
    for n, p in model.named_parameters():
        if "adapter_2" in n:
            p.requires_grad = False
    wrapped = tr.accelerator.prepare(model)
    opt = tr.create_optimizer(wrapped)
    sched = get_polynomial_decay_schedule_with_warmup(opt, 2, 20, lr_end=0, power=1)
    wrapped.train()
    return model, tr, wrapped, opt, sched


batches = [to_device(make_vilt_batch(2, 16, 224, 100, seed=50 + i), "cuda") for i in range(5)]
m1, tr1, w1, o1, s1 = build()
eager, lrs = [], []
for i in range(5):
    eager.append(tr1.train_step(w1, i, batches[i], o1, s1).item())
    lrs.append([float(g["lr"]) for g in o1.param_groups])
    print("eager", i, eager[-1], [x.item() for x in tr1.last_objectives], lrs[-1], s1.last_epoch)
m2, tr2, w2, o2, s2 = build()
g = GraphedTrainStep(tr2, w2, o2, s2, batches[0], warmup=int(sys.argv[1]) if len(sys.argv) > 1 else 2)
for i in range(5):
    v = g(batches[i]).item()
    print("graph", i, v, [x.item() for x in tr2.last_objectives], [float(q["lr"]) for q in o2.param_groups],
          s2.last_epoch, g.lr_buf.tolist())
worst = []
for (n1, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
    if "adapter_0" in n1 or "adapter_1" in n1 or "task_layer" in n1:
        worst.append(((p1 - p2).abs().max().item(), n1))
print(sorted(worst)[-5:])
for k, st in list(o2.state.items())[:3]:
    print("state step", st["step"])
