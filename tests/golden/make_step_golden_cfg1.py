"""Step-level golden at the BENCHMARKED configuration's shapes (BASELINE configs[1]: ViLT-B/32 + DAT
rank 128, 384x384 image / 40 text tokens, MKD temperature 2.0; batch 4 instead of 32 so the CPU run
finishes in minutes), produced by the REFERENCE's own ``TaskTrainer.train_step`` / ``create_optimizer`` /
``kl_loss`` (imported unmodified from /root/reference) driving the REFERENCE's ``Adapter`` inside HF
``ViltModel`` -- twice:

  fp32/...   plain fp32 on the CPU
  bf16/...   the forward under ``torch.autocast('cpu', dtype=torch.bfloat16)`` with fp32 outputs, which is
             what the reference's launcher does through accelerate (mixed precision + ConvertOutputsToFp32;
             accelerate_config.yaml:8 says fp16, SURVEY.md section 5 sets bf16 as the comparison point)

Recorded per step: the three logits of the MKD schedule, loss_0, and -- PRE-Adam -- the gradient of every
trainable tensor after the backward of pass B and of pass C, as its norm plus a seeded 8-column random
sketch (``tests/golden_inputs.py::grad_sketch``; biases in full), so the fixture stays ~2 MB.

    python tests/golden/make_step_golden_cfg1.py     # writes tests/golden/step_golden_cfg1.npz
"""
from __future__ import annotations

import sys
import time
from argparse import Namespace
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))

import make_step_golden as base  # noqa: E402
from tests.golden_inputs import grad_sketch  # noqa: E402

SEED, RANK, STEPS, MAX_STEPS, LR, TEMP = 4321, 128, 2, 10, 1e-4, 2.0
B, T, H, C = 4, 40, 384, 100
TASK = "art"


def run(variant, sd0, gold):
    from feddat_b200.synthetic import make_vilt_batch
    base.RANK, base.C, base.TASK = RANK, C, TASK
    ref = base.build_reference_model(sd0)
    for p in ref.parameters():                                     # main.py:138-139
        p.requires_grad = False
    for n, p in ref.named_parameters():                            # main.py:157-159, 248-250
        if "adapter" in n or "task" in n:
            p.requires_grad = True
    for n, p in ref.named_parameters():                            # task_trainer.py:43-45
        if "adapter_2" in n:
            p.requires_grad = False
    sd = ref.state_dict()                                          # task_trainer.py:36-41
    for name in sd:
        if "adapter_1" in name:
            sd[name.replace("adapter_1", "adapter_2")].data.copy_(sd[name].data)

    sys.path.insert(0, "/root/reference")
    import src.train.visionlanguage_tasks.task_trainer as ref_tt   # the reference's own trainer
    from transformers import get_polynomial_decay_schedule_with_warmup

    class Wrap(nn.Module):                                         # stands in for accelerate's prepared model
        def __init__(self, m):
            super().__init__()
            self.module = m

        def forward(self, *a, **k):
            if variant == "bf16":
                with torch.autocast("cpu", dtype=torch.bfloat16):
                    out = self.module(*a, **k)
                return tuple(o.float() for o in out)               # accelerate: ConvertOutputsToFp32
            return self.module(*a, **k)

    grads = []

    def backward(loss):
        loss.backward()
        grads.append({n: p.grad.detach().clone() for n, p in ref.named_parameters() if p.grad is not None})

    tr = ref_tt.TaskTrainer()
    tr.args = Namespace(optimizer_mode="dat", encoder_name="vilt")
    tr.accelerator = Namespace(device=torch.device("cpu"), backward=backward)
    tr.device = torch.device("cpu")
    tr.task_key = TASK
    tr.batch2inputs_converter = lambda b: dict(b["encodings"])
    tr.loss_criterion = nn.BCEWithLogitsLoss(reduction="mean")     # train_vqa_crossvqa.py:237
    tr.weight_decay, tr.lr, tr.adam_epsilon = 1e-2, LR, 1e-8
    # BASELINE configs[1] uses temperature 2.0; kl_loss's default is 3 and train_step passes none
    tr.kl_criterion = lambda out, tgt: ref_tt.kl_loss(out, tgt, temp=TEMP)
    wrapped = Wrap(ref)
    opt = tr.create_optimizer(wrapped)
    sched = get_polynomial_decay_schedule_with_warmup(opt, num_warmup_steps=int(MAX_STEPS * 0.1),
                                                      num_training_steps=MAX_STEPS, lr_end=0, power=1)
    wrapped.train()
    for step in range(STEPS):
        batch = make_vilt_batch(B, T, H, C, seed=SEED + step)
        seen = []
        h = ref.task_layer[TASK].register_forward_hook(lambda m, i, o: seen.append(o.detach().float().clone()))
        grads.clear()
        t0 = time.time()
        loss_0 = tr.train_step(wrapped, step, batch, opt, sched)
        h.remove()
        gold[f"{variant}/step{step}/loss_0"] = np.array(loss_0.item())
        for name, t in zip(("logits_all", "logits_1", "logits_0"), seen):
            gold[f"{variant}/step{step}/{name}"] = t.numpy()
        assert len(grads) == 2
        for tag, gd in zip(("B", "C"), grads):
            for n, g in gd.items():
                gold[f"{variant}/step{step}/grad{tag}/norm/{n}"] = np.array(g.double().norm().item())
                gold[f"{variant}/step{step}/grad{tag}/sketch/{n}"] = grad_sketch(n, g.float().numpy())
        print(f"[{variant}] step {step}: loss_0 = {loss_0.item():.6f}  ({time.time() - t0:.1f} s; "
              f"{len(grads[0])} / {len(grads[1])} gradient tensors after pass B / C)")
    gold[f"{variant}/n_optimizer_tensors"] = np.array(sum(len(g["params"]) for g in opt.param_groups))


def main():
    from feddat_b200.train.prepare import default_args, prepare_model
    torch.manual_seed(SEED)
    ours = prepare_model(default_args(ordered_cl_tasks=[TASK], adapter_rank=RANK), place=False)
    sd0 = {k: v.clone() for k, v in ours.state_dict().items()}
    gold = {"meta": np.array([SEED, RANK, STEPS, MAX_STEPS, B, T, H, C]), "lr": np.array(LR), "temp": np.array(TEMP)}
    for variant in ("fp32", "bf16"):
        run(variant, sd0, gold)
    np.savez_compressed(ROOT / "tests" / "golden" / "step_golden_cfg1.npz", **gold)
    print("wrote step_golden_cfg1.npz with", len(gold), "arrays")


if __name__ == "__main__":
    main()
