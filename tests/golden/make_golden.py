"""Generates the op-level golden vectors by EXECUTING THE REFERENCE'S OWN CODE (run in the build
container, where /root/reference is mounted; the GPU box only sees the committed .npz files).

    python tests/golden/make_golden.py

What is executed from the reference (nothing is copied into this repo):
  * class Adapter          src/modeling/models/adapter.py (exec'd from source; the two hard-coded
                           ``.to('cuda')`` at :144,160 are redirected to the input's device so the
                           gating branch runs on CPU -- SURVEY.md F6)
  * kl_loss                src/train/visionlanguage_tasks/task_trainer.py:506-516 (imported)
  * get_average_net        src/train/main.py:50-65 (function source exec'd on its own: main.py as a
                           module imports accelerate / adapter-transformers, absent here)
Inputs come from numpy's PCG64 with fixed seeds, so tests can regenerate them bit-exactly.
"""
from __future__ import annotations

import ast
import contextlib
import io
import sys
import types
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def load_reference_adapter():
    src = (REF / "src/modeling/models/adapter.py").read_text()
    assert src.count(".to('cuda')") == 2
    src = src.replace(".to('cuda')", ".to(hidden_states.device)")
    mod = types.ModuleType("ref_adapter")
    exec(compile(src, str(REF / "src/modeling/models/adapter.py"), "exec"), mod.__dict__)
    return mod.Adapter


def load_reference_kl_loss():
    sys.path.insert(0, str(REF))
    from src.train.visionlanguage_tasks.task_trainer import kl_loss  # noqa: PLC0415
    return kl_loss


def load_reference_get_average_net():
    path = REF / "src/train/main.py"
    tree = ast.parse(path.read_text())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "get_average_net")
    mod = types.ModuleType("ref_main_fn")
    mod.__dict__["torch"] = torch
    exec(compile(ast.Module(body=[fn], type_ignores=[]), str(path), "exec"), mod.__dict__)
    return mod.get_average_net


def rng_weights(rng, r, d=768, names=("adapter_0", "adapter_1", "adapter_2")):
    """Deterministic adapter weights (std 0.05 so the bottleneck output is not negligible; non-zero
    biases so bias paths are exercised -- the reference's init is N(0, .02)/zero, adapter.py:5-14)."""
    w = {}
    for n in names:
        w[f"{n}_down.weight"] = (rng.standard_normal((r, d)) * 0.05).astype(np.float32)
        w[f"{n}_down.bias"] = (rng.standard_normal((r,)) * 0.1).astype(np.float32)
        w[f"{n}_up.weight"] = (rng.standard_normal((d, r)) * 0.05).astype(np.float32)
        w[f"{n}_up.bias"] = (rng.standard_normal((d,)) * 0.1).astype(np.float32)
    return w


def adapter_case(Adapter, seed, r, shape, mode, dtype=torch.float32):
    d = 768
    assert d % r == 0
    rng = np.random.default_rng(seed)
    w = rng_weights(rng, r)
    x = rng.standard_normal(shape + (d,)).astype(np.float32)
    g = rng.standard_normal(shape + (d,)).astype(np.float32)
    with contextlib.redirect_stdout(io.StringIO()):   # the reference prints the names (:27)
        ad = Adapter(names=["adapter_0", "adapter_1", "adapter_2"], device="cpu", model_dim=d,
                     adapter_reduction_factor=d // r)
    ad.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
    ad = ad.to(dtype)
    if mode == "single":            # task_trainer.py:290-291
        ad.deactivate_gating()
        ad.set_active_adapter("adapter_1")
        train = "adapter_1"
    else:                           # task_trainer.py:311-312
        ad.activate_gating()
        ad.set_active_adapter("adapter_0")
        train = "adapter_0"
    xt = torch.from_numpy(x).to(dtype).requires_grad_(True)
    y = ad(xt, xt)                  # adaptered_output.py:78: adapter(h, h)
    y.backward(torch.from_numpy(g).to(dtype))
    out = {"y": y.detach().float().numpy(), "dx": xt.grad.float().numpy()}
    for part in ("down", "up"):
        lin = getattr(ad, f"{train}_{part}")
        out[f"d_{part}_w"] = lin.weight.grad.float().numpy()
        out[f"d_{part}_b"] = lin.bias.grad.float().numpy()
    # requires_grad pattern after the mode switch (adapter.py:71-85, :55-58)
    out["requires_grad"] = np.array([int(p.requires_grad) for _, p in sorted(ad.named_parameters())])
    return out


def bert_case(Adapter, seed, r, shape):
    d = 768
    rng = np.random.default_rng(seed)
    w = rng_weights(rng, r)
    ffn = rng.standard_normal(shape + (d,)).astype(np.float32)
    x = rng.standard_normal(shape + (d,)).astype(np.float32)
    lnw = (1.0 + 0.1 * rng.standard_normal(d)).astype(np.float32)
    lnb = (0.1 * rng.standard_normal(d)).astype(np.float32)
    with contextlib.redirect_stdout(io.StringIO()):
        ad = Adapter(names=["adapter_0", "adapter_1", "adapter_2"], device="cpu", model_dim=d,
                     adapter_reduction_factor=d // r)
    ad.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
    ln = nn.LayerNorm(d, eps=1e-12)
    ln.weight.data.copy_(torch.from_numpy(lnw))
    ln.bias.data.copy_(torch.from_numpy(lnb))
    ad.activate_gating()
    ad.set_active_adapter("adapter_0")
    y = ad.adapter_layer_forward_bert(torch.from_numpy(ffn), torch.from_numpy(x), ln)
    return {"y": y.detach().numpy()}


def main():
    Adapter = load_reference_adapter()
    kl_loss = load_reference_kl_loss()
    get_average_net = load_reference_get_average_net()
    gold = {}

    for name, seed, r, shape, mode in [
        ("single_r16", 11, 16, (2, 7), "single"),
        ("gating_r16", 12, 16, (2, 7), "gating"),
        ("single_r48", 13, 48, (2, 33), "single"),
        ("gating_r48", 14, 48, (2, 33), "gating"),
        ("gating_r128", 15, 128, (1, 70), "gating"),
    ]:
        res = adapter_case(Adapter, seed, r, shape, mode)
        if r >= 48:  # keep fixtures small: the weight grads of the wide cases stay out
            res = {k: v for k, v in res.items() if k in ("y", "dx", "d_down_b", "d_up_b", "requires_grad")}
        for k, v in res.items():
            gold[f"adapter/{name}/{k}"] = v
        gold[f"adapter/{name}/meta"] = np.array([seed, r, *shape, int(mode == "gating")])

    res = bert_case(Adapter, 21, 16, (2, 5))
    gold["bert/gating_r16/y"] = res["y"]
    gold["bert/gating_r16/meta"] = np.array([21, 16, 2, 5, 1])

    # kl_loss (+ BCE task loss and the (a+b)/2 combination of task_trainer.py:299-301)
    for name, seed, shape, temp in [("vilt_T3", 31, (4, 100), 3.0), ("vilt_T2", 32, (32, 100), 2.0),
                                    ("wide_T3", 33, (2, 3, 3001), 3.0)]:
        rng = np.random.default_rng(seed)
        a = (rng.standard_normal(shape) * 2).astype(np.float32)
        b = (rng.standard_normal(shape) * 2).astype(np.float32)
        at = torch.from_numpy(a).requires_grad_(True)
        loss = kl_loss(at, torch.from_numpy(b), temp=temp)
        loss.backward()
        gold[f"kl/{name}/loss"] = np.array(loss.item(), np.float64)
        gold[f"kl/{name}/grad"] = at.grad.numpy()
        gold[f"kl/{name}/meta"] = np.array([seed, temp, *shape])
        if len(shape) == 2:
            tgt = np.zeros(shape, np.float32)
            for i in range(shape[0]):
                for j in rng.choice(shape[1], size=rng.integers(1, 4), replace=False):
                    tgt[i, j] = rng.choice([0.3, 0.6, 0.9, 1.0])
            at2 = torch.from_numpy(a).requires_grad_(True)
            crit = nn.BCEWithLogitsLoss(reduction="mean")          # train_vqa_crossvqa.py:237
            task = crit(at2, torch.from_numpy(tgt)) * tgt.shape[1]  # task_trainer.py:299
            total = (task + kl_loss(at2, torch.from_numpy(b).clone().detach(), temp=temp)) / 2
            total.backward()
            gold[f"mkd/{name}/target"] = tgt
            gold[f"mkd/{name}/task"] = np.array(task.item(), np.float64)
            gold[f"mkd/{name}/total"] = np.array(total.item(), np.float64)
            gold[f"mkd/{name}/grad"] = at2.grad.numpy()

    # get_average_net
    class Server:
        def __init__(self, sd):
            self._sd = sd
            self.comm_state_dict_names = list(sd.keys())

        def state_dict(self):
            return self._sd

    for name, seed, nums in [("equal3", 41, [1, 1, 1]), ("weighted3", 42, [3, 1, 2]),
                             ("equal8", 43, [1] * 8)]:
        rng = np.random.default_rng(seed)
        keys = ["a.adapter_1_down.weight", "a.adapter_1_up.bias"]
        clients = [{k: torch.from_numpy(rng.standard_normal(257 if "bias" in k else (16, 33)).astype(np.float32))
                    for k in keys} for _ in nums]
        server = Server({k: torch.zeros_like(v) for k, v in clients[0].items()})
        get_average_net(server, clients, nums, None, "cpu")
        for k in keys:
            gold[f"fedavg/{name}/{k}"] = server.state_dict()[k].numpy()
        gold[f"fedavg/{name}/meta"] = np.array([seed, *nums])

    np.savez_compressed(OUT / "op_golden.npz", **gold)
    total = sum(v.nbytes for v in gold.values())
    print(f"wrote {OUT / 'op_golden.npz'}: {len(gold)} arrays, {total / 1e6:.2f} MB raw")


if __name__ == "__main__":
    main()
