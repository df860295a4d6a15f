"""Step-level golden: the REFERENCE's own ``TaskTrainer.train_step`` / ``create_optimizer`` /
``kl_loss`` (imported unmodified from /root/reference) driving the REFERENCE's ``Adapter`` (exec'd
from source, '.to("cuda")' redirected for CPU) inside HF ``ViltModel``, fp32 on CPU, on the
BASELINE config-0 batch (B=2, 224x224, 32 tokens, rank 64).  Recipe: SURVEY.md Appendix A.1.

    python tests/golden/make_step_golden.py        # writes tests/golden/step_golden.npz

The initial weights come from this repo's ``prepare_model`` under torch.manual_seed(SEED) on CPU --
the state-dict keys are identical to the reference's, so the same state dict loads into both sides.
"""
from __future__ import annotations

import contextlib
import io
import sys
from argparse import Namespace
from collections import OrderedDict
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))

from make_golden import load_reference_adapter  # noqa: E402

SEED, RANK, STEPS, MAX_STEPS, LR = 1234, 64, 3, 10, 1e-4
B, T, H, C = 2, 32, 224, 100
TASK = "art"


def build_reference_model(state_dict):
    """HF ViltModel + the reference Adapter injected through a restated Adaptered_ViltOutput
    (adaptered_output.py:67-79) + the reference's head layout (vilt.py:202-209)."""
    from transformers import ViltConfig, ViltModel
    RefAdapter = load_reference_adapter()

    class RefAdapteredViltOutput(nn.Module):
        def __init__(self, layer):
            super().__init__()
            self.layer = layer
            with contextlib.redirect_stdout(io.StringIO()):
                self.adapter = RefAdapter(names=["adapter_0", "adapter_1", "adapter_2"], device="cpu",
                                          model_dim=768, adapter_reduction_factor=768 // RANK)

        def forward(self, hidden_states, input_tensor):
            hidden_states = self.layer.dense(hidden_states)
            hidden_states = self.layer.dropout(hidden_states)
            hidden_states = hidden_states + input_tensor
            return self.adapter(hidden_states, hidden_states)

    class Enc(nn.Module):
        def __init__(self):
            super().__init__()
            self.vilt = ViltModel(ViltConfig())
            emb = self.vilt.embeddings.token_type_embeddings.weight.data      # vilt.py:102-113
            self.vilt.embeddings.token_type_embeddings = nn.Embedding(3, 768)
            self.vilt.embeddings.token_type_embeddings.weight.data[:2] = emb[:2]

    class RefLearner(nn.Module):
        def __init__(self):
            super().__init__()
            self.vilt_encoder = Enc()
            for i in range(12):
                self.vilt_encoder.vilt.encoder.layer[i].output = RefAdapteredViltOutput(
                    self.vilt_encoder.vilt.encoder.layer[i].output)
            self.task_layer = nn.ModuleDict({TASK: nn.Sequential(OrderedDict([
                ("clf_fc0", nn.Linear(768, 1536)), ("clf_norm0", nn.LayerNorm(1536)),
                ("clf_actv0", nn.GELU()), ("clf_fc1", nn.Linear(1536, C))]))})

        def _adapters(self):
            return [l.output.adapter for l in self.vilt_encoder.vilt.encoder.layer]

        def set_active_adapter(self, name):          # vilt.py:363-365
            [a.set_active_adapter(name) for a in self._adapters()]

        def activate_gating(self):                   # vilt.py:367-369
            [a.activate_gating() for a in self._adapters()]

        def deactivate_gating(self):                 # vilt.py:371-373
            [a.deactivate_gating() for a in self._adapters()]

        def forward(self, task_key, **enc):
            enc.pop("dense_masks", None)
            pooled = self.vilt_encoder.vilt(**enc).pooler_output
            return pooled, self.task_layer[task_key](pooled)

    m = RefLearner()
    missing, unexpected = m.load_state_dict(state_dict, strict=False)
    assert not unexpected, unexpected
    assert all("position_ids" in k or "token_type_ids" in k for k in missing), missing
    return m


def main():
    from feddat_b200.synthetic import make_vilt_batch
    from feddat_b200.train.prepare import default_args, prepare_model

    torch.manual_seed(SEED)
    args = default_args(ordered_cl_tasks=[TASK], adapter_rank=RANK)
    ours = prepare_model(args, place=False)
    sd0 = {k: v.clone() for k, v in ours.state_dict().items()}
    # non-zero adapter biases / wider weights would hide nothing here: keep the reference init
    # (N(0, .02), zero bias: adapter.py:5-14) -- this is what a real first round starts from

    ref = build_reference_model(sd0)
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
    from src.train.visionlanguage_tasks.task_trainer import TaskTrainer  # the reference's own trainer
    from transformers import get_polynomial_decay_schedule_with_warmup

    class Wrap(nn.Module):                                         # stands in for DDP (.module)
        def __init__(self, m):
            super().__init__()
            self.module = m

        def forward(self, *a, **k):
            return self.module(*a, **k)

    tr = TaskTrainer()
    tr.args = Namespace(optimizer_mode="dat", encoder_name="vilt")
    tr.accelerator = Namespace(device=torch.device("cpu"), backward=lambda l: l.backward())
    tr.device = torch.device("cpu")
    tr.task_key = TASK
    tr.batch2inputs_converter = lambda b: dict(b["encodings"])
    tr.loss_criterion = nn.BCEWithLogitsLoss(reduction="mean")     # train_vqa_crossvqa.py:237
    tr.weight_decay, tr.lr, tr.adam_epsilon = 1e-2, LR, 1e-8
    wrapped = Wrap(ref)
    opt = tr.create_optimizer(wrapped)
    sched = get_polynomial_decay_schedule_with_warmup(opt, num_warmup_steps=int(MAX_STEPS * 0.1),
                                                      num_training_steps=MAX_STEPS, lr_end=0, power=1)
    n_opt = sum(len(g["params"]) for g in opt.param_groups)

    gold = {"meta": np.array([SEED, RANK, STEPS, MAX_STEPS, B, T, H, C]), "lr": np.array(LR),
            "n_optimizer_tensors": np.array(n_opt)}
    wrapped.train()
    for step in range(STEPS):
        batch = make_vilt_batch(B, T, H, C, seed=SEED + step)
        # capture the three logits of the MKD schedule through forward hooks on the head
        seen = []
        h = ref.task_layer[TASK].register_forward_hook(lambda m, i, o: seen.append(o.detach().clone()))
        loss_0 = tr.train_step(wrapped, step, batch, opt, sched)
        h.remove()
        gold[f"step{step}/loss_0"] = np.array(loss_0.item())
        for name, t in zip(("logits_all", "logits_1", "logits_0"), seen):
            gold[f"step{step}/{name}"] = t.numpy()
        print(f"step {step}: loss_0 = {loss_0.item():.6f}")
    sd1 = ref.state_dict()
    for k in sd1:
        if ("adapter_0" in k or "adapter_1" in k or "task_layer" in k):
            gold[f"delta_norm/{k}"] = np.array((sd1[k] - sd0[k]).double().norm().item())
            gold[f"final_norm/{k}"] = np.array(sd1[k].double().norm().item())
    np.savez_compressed(ROOT / "tests" / "golden" / "step_golden.npz", **gold)
    print("optimizer tensors:", n_opt, "; wrote step_golden.npz with", len(gold), "arrays")


if __name__ == "__main__":
    main()
