"""Debug aid: graph-captured fwd+bwd(+opt) of pass B vs eager, per-parameter gradient diff."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import debug_graph as dg  # noqa: E402
from feddat_b200 import ops  # noqa: E402
from feddat_b200.train.task_trainer import mkd_objective  # noqa: E402

batches = dg.batches
flags = []
_orig = ops.dat_backward
def _spy(*a, **k):
    flags.append(torch.cuda.is_current_stream_capturing())
    return _orig(*a, **k)
ops.dat_backward = _spy
import feddat_b200.modeling.adapter as ad
ad.ops.dat_backward = _spy


def pass_b(m, batch, teacher, opt=None):
    m.deactivate_gating(); m.set_active_adapter("adapter_1")
    _, logits = m(task_key="art", **dict(batch["encodings"]))
    L, _ = mkd_objective(logits, teacher, batch["target_scores"], 2.0)
    L.backward()
    if opt is not None:
        opt.step()
    return L


def clear(m):
    m.vilt_encoder._embed_cache = None
    for a in m._adapters():
        a._pack_cache.clear()


for with_opt in (False, True):
    res = {}
    for mode in ("eager", "graph"):
        m, tr, w, o, s = dg.build()
        teacher = torch.randn(2, 100, device="cuda")
        b = batches[0]
        # one eager step first so optimizer state exists (as in GraphedTrainStep warm-up)
        pass_b(m, b, teacher, o); o.zero_grad()
        if mode == "eager":
            L = pass_b(m, b, teacher, o if with_opt else None)
        else:
            clear(m); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            flags.clear()
            with torch.cuda.graph(g):
                L = pass_b(m, b, teacher, o if with_opt else None)
            print("capturing flags in backward:", set(flags), len(flags))
            g.replay(); torch.cuda.synchronize()
        res[mode] = (L.item(), {n: (p.grad.clone() if p.grad is not None else None) for n, p in m.named_parameters()},
                     {n: p.detach().clone() for n, p in m.named_parameters() if p.requires_grad or "adapter" in n or "task" in n})
    print("with_opt", with_opt, "L eager/graph", res["eager"][0], res["graph"][0])
    worst = []
    for n, ge in res["eager"][1].items():
        gg = res["graph"][1][n]
        if (ge is None) != (gg is None):
            print("grad presence differs:", n, ge is None, gg is None); continue
        if ge is not None:
            worst.append(((ge - gg).abs().max().item() / (ge.abs().max().item() + 1e-20), n))
    print("worst rel grad diffs:", sorted(worst)[-4:])
    pw = [((res["eager"][2][n] - res["graph"][2][n]).abs().max().item(), n) for n in res["eager"][2]]
    print("worst param diffs:", sorted(pw)[-4:])
