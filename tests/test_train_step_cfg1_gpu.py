"""Step-level parity at the BENCHMARKED configuration's shapes (BASELINE configs[1]: rank 128, 384x384 / 40
tokens, temperature 2; batch 4): this repo's ``TaskTrainer.train_step`` -- both the reference pass order and
the batched schedule -- against goldens from the REFERENCE's own trainer run (a) in fp32 and (b) under bf16
autocast (tests/golden/make_step_golden_cfg1.py).  Compared: the three logits of the MKD schedule, loss_0,
and the PRE-Adam gradients of pass B (adapter_1 + head) and pass C (adapter_0 + head) through norms and
seeded random sketches -- gradients, not post-Adam trajectories (Adam's first steps are sign-like and
amplify rounding noise into +-lr moves).  GPU only."""
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.nn as nn

from tests.golden_inputs import grad_sketch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    return np.load(Path(__file__).resolve().parent / "golden" / "step_golden_cfg1.npz")


def fro(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def build(gold, batched):
    from feddat_b200.modeling.vilt import convert_batch_to_vilt_input_dict
    from feddat_b200.train.accelerator import Accelerator
    from feddat_b200.train.prepare import default_args, place_on_gpu, prepare_model
    from feddat_b200.train.task_trainer import TaskTrainer, get_polynomial_decay_schedule_with_warmup
    seed, rank, steps, max_steps, B, T, H, C = (int(v) for v in gold["meta"])
    torch.manual_seed(seed)
    model = prepare_model(default_args(ordered_cl_tasks=["art"], adapter_rank=rank), place=False)
    place_on_gpu(model)
    tr = TaskTrainer()
    tr.batched_passes = batched
    tr.args = SimpleNamespace(optimizer_mode="dat", encoder_name="vilt", debug=0)
    tr.accelerator = Accelerator(device="cuda")
    tr.device, tr.task_key = torch.device("cuda"), "art"
    tr.batch2inputs_converter = convert_batch_to_vilt_input_dict
    tr.loss_criterion = nn.BCEWithLogitsLoss(reduction="mean")
    tr.weight_decay, tr.lr, tr.adam_epsilon, tr.kl_temp = 1e-2, float(gold["lr"]), 1e-8, float(gold["temp"])
    sd = model.state_dict()
    for name in sd:                                                 # task_trainer.py:36-45
        if "adapter_1" in name:
            sd[name.replace("adapter_1", "adapter_2")].data.copy_(sd[name].data)
    for n, p in model.named_parameters():
        if "adapter_2" in n:
            p.requires_grad = False
    wrapped = tr.accelerator.prepare(model)
    opt = tr.create_optimizer(wrapped)
    sched = get_polynomial_decay_schedule_with_warmup(opt, int(max_steps * 0.1), max_steps, lr_end=0, power=1)
    wrapped.train()
    return model, tr, wrapped, opt, sched, (seed, steps, B, T, H, C)


@pytest.mark.parametrize("batched", [False, True], ids=["reference_order", "batched_schedule"])
def test_cfg1_step_logits_and_gradients(gold, batched):
    from feddat_b200.synthetic import make_vilt_batch, to_device
    model, tr, wrapped, opt, sched, (seed, steps, B, T, H, C) = build(gold, batched)
    assert sum(len(g["params"]) for g in opt.param_groups) == int(gold["bf16/n_optimizer_tensors"])
    probes = {}

    def probe(tag, _m):
        probes[tag] = {n: p.grad.detach().float().cpu().numpy() for n, p in model.named_parameters()
                       if p.grad is not None}

    tr.grad_probe = probe
    rep = []
    for step in range(steps):
        probes.clear()
        batch = to_device(make_vilt_batch(B, T, H, C, seed=seed + step), "cuda")
        loss_0 = tr.train_step(wrapped, step, batch, opt, sched)
        torch.cuda.synchronize()
        assert ("BC" in probes) == batched, "the expected schedule did not run"
        got = {n: t.float().cpu().numpy() for n, t in zip(("logits_all", "logits_1", "logits_0"), tr.last_logits)}
        for variant in ("bf16", "fp32"):
            for n in got:
                rep.append((variant, step, n, fro(got[n], gold[f"{variant}/step{step}/{n}"])))
            w = float(gold[f"{variant}/step{step}/loss_0"])
            rep.append((variant, step, "loss_0", abs(loss_0.item() - w) / w))
        # gradients: pass B = adapter_1 (+ head), pass C = adapter_0 (+ head)
        if batched:
            passes = {"B": {n: g for n, g in probes["BC"].items() if "adapter_1" in n},
                      "C": {n: g for n, g in probes["BC"].items() if "adapter_1" not in n}}
            passes["B"].update({n: g for n, g in probes["B_head"].items() if "task_layer" in n})
        else:
            passes = {"B": probes["B"], "C": probes["C"]}
        for tag, gd in passes.items():
            want_names = {k.split("/sketch/")[1] for k in gold.files if k.startswith(f"bf16/step{step}/grad{tag}/sketch/")}
            assert set(gd) == want_names, (tag, sorted(set(gd) ^ want_names)[:4])
            for variant in ("bf16", "fp32"):
                num = den = 0.0
                worst = (0.0, "")
                for n, g in gd.items():
                    want = gold[f"{variant}/step{step}/grad{tag}/sketch/{n}"].astype(np.float64)
                    d = np.linalg.norm(grad_sketch(n, g).astype(np.float64) - want) ** 2
                    num += d; den += np.linalg.norm(want) ** 2
                    e = (d ** 0.5) / max(np.linalg.norm(want), 1e-30)
                    worst = max(worst, (e, n))
                rep.append((variant, step, f"grad{tag} (all tensors)", (num / den) ** 0.5))
                rep.append((variant, step, f"grad{tag} worst tensor {worst[1][-40:]}", worst[0]))
    print(f"\ncfg1-shape step parity ({'batched' if batched else 'reference-order'} schedule), relative Frobenius:")
    for variant, step, what, e in rep:
        print(f"  vs {variant} reference  step{step}  {what:62s} {e:.4f}")
    # Step 0 is pure forward / backward parity (the first optimizer step of a schedule has lr = 0, so no
    # parameter has moved when pass C runs).  Measured on B200 (profiles/r2_summary.md): logits 1.3-1.5e-2 from
    # BOTH goldens, loss_0 1e-4, whole-pass gradients 1.0e-2 (bf16 golden) / 1.2e-2 (fp32 golden).  For scale: the
    # reference's own bf16-autocast run sits 7.3e-3 (logits) / 6.3e-3 (gradients) from its own fp32 run.
    # autocast rounds only the Linear operands and keeps LayerNorm / residual adds in fp32; this path also
    # keeps the residual stream, the LayerNorm outputs and the bottleneck's hidden in bf16 (half the HBM
    # bytes of every elementwise pass), which is where the second 7e-3 comes from.  Bars: 2e-2 on logits
    # (12 layers of bf16 rounding; the op-level tests hold the north-star 1e-2 per operator), 1e-3 on the
    # loss, 2e-2 on whole-pass gradients.
    for variant, step, what, e in rep:
        if step != 0:
            continue
        if what.startswith("logits"):
            assert e < 2e-2, (variant, what, e)
        elif what == "loss_0":
            assert e < 1e-3, (variant, what, e)
        elif "(all tensors)" in what:
            assert e < 2e-2, (variant, what, e)
    # Step 1 carries one AdamW update (sign-like): bars as in tests/test_train_step_gpu.py
    for variant, step, what, e in rep:
        if step == 1 and (what.startswith("logits") or what == "loss_0" or "(all tensors)" in what):
            assert e < 4e-2, (variant, what, e)
