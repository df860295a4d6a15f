"""Debug aid: one train_step from identical state, eager vs graph-captured: per-pass logits + param diffs."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import debug_graph as dg  # noqa: E402
from feddat_b200.train.graphed import GraphedTrainStep, _ReplayScheduler  # noqa: E402

batches = dg.batches
for reuse in (True, False):
    out = {}
    for mode in ("eager", "graph"):
        m, tr, w, o, s = dg.build()
        tr.reuse_gating_forward = reuse
        for i in range(2):
            tr.train_step(w, i, batches[i], o, s)
        if mode == "eager":
            loss = tr.train_step(w, 2, batches[2], o, s)
        else:
            g = GraphedTrainStep(tr, w, o, s, batches[2], warmup=0)
            loss = g(batches[2])
        torch.cuda.synchronize()
        out[mode] = (loss.item(), [t.clone() for t in tr.last_logits], [t.item() for t in tr.last_objectives],
                     {n: p.detach().clone() for n, p in m.named_parameters() if "adapter" in n or "task" in n})
    e, g_ = out["eager"], out["graph"]
    print("reuse", reuse, "loss", e[0], g_[0], "objectives", e[2], g_[2])
    print("  logits max diff A/B/C:", [(a - b).abs().max().item() for a, b in zip(e[1], g_[1])])
    pw = sorted(((e[3][n] - g_[3][n]).abs().max().item(), n) for n in e[3])
    print("  worst param diffs:", pw[-3:])
    grp = {}
    for d, n in pw:
        k = "adapter_0" if "adapter_0" in n else "adapter_1" if "adapter_1" in n else "adapter_2" if "adapter_2" in n else "head"
        grp[k] = max(grp.get(k, 0.0), d)
    print("  by group:", grp)
