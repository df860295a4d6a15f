"""Debug aid: is a graph-captured forward/step numerically identical to eager? and is eager run-to-run deterministic?"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from debug_graph import build, batches  # noqa: E402  (runs the 5-step comparison too)

# 1) eager run-to-run
outs = []
for rep in range(2):
    m, tr, w, o, s = build()
    ls = [tr.train_step(w, i, batches[i], o, s).item() for i in range(3)]
    outs.append(ls)
print("eager run-to-run", outs)

# 2) forward only: eager vs captured
m, tr, w, o, s = build()
m.activate_gating(); m.set_active_adapter("adapter_0")
enc = dict(batches[0]["encodings"])
with torch.no_grad():
    _, l_e1 = m(task_key="art", **dict(enc))
    _, l_e2 = m(task_key="art", **dict(enc))
print("eager fwd repeat max diff", (l_e1 - l_e2).abs().max().item(), "logit absmax", l_e1.abs().max().item())
m.vilt_encoder._embed_cache = None
for a in m._adapters():
    a._pack_cache.clear()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.no_grad(), torch.cuda.graph(g):
    _, l_g = m(task_key="art", **dict(enc))
g.replay(); torch.cuda.synchronize()
print("graph fwd vs eager max diff", (l_g - l_e1).abs().max().item())
g.replay(); torch.cuda.synchronize()
print("graph fwd replay2 vs eager max diff", (l_g - l_e1).abs().max().item())
