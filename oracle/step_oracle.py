"""Oracle: the reference's per-client train step as plain PyTorch on the CPU.  TEST / BASELINE ONLY.

A restatement ("port") of the reference path that cannot travel to the GPU box: the DAT ``Adapter``
(src/modeling/models/adapter.py), its ViLT injection wrapper (src/modeling/adaptered_output.py:67-79),
the task head (src/modeling/vilt.py:202-209) and the ``dat`` branch of ``TaskTrainer.train_step``
with ``create_optimizer`` and ``kl_loss`` (src/train/visionlanguage_tasks/task_trainer.py:280-330,
477-516), around the SAME HF ``ViltModel`` the product uses.  Pinned by tests/test_step_oracle.py
against tests/golden/step_golden.npz (produced by the reference's own TaskTrainer + Adapter).

Used by: tests, ``bench.py``'s ``cpu_baseline`` leg and ``bench.py --impl reference``.
Never imported by ``feddat_b200``.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.optim import AdamW


class OracleAdapter(nn.Module):
    """adapter.py:16-163 as eager PyTorch (the arithmetic the CUDA kernels replace)."""

    def __init__(self, names, rank, model_dim=768):
        super().__init__()
        self.gating = False
        self.scaling = 1.0                                           # adapter.py:25
        for name in names:
            setattr(self, f"{name}_down", nn.Linear(model_dim, rank))
            setattr(self, f"{name}_up", nn.Linear(rank, model_dim))
        for m in self.modules():
            if isinstance(m, nn.Linear):                             # adapter.py:5-14
                m.weight.data.normal_(0.0, 0.02)
                m.bias.data.zero_()
        self._active = None

    def set_active_adapter(self, name):                              # adapter.py:66-95
        self._active = name
        other = {"adapter_0": "adapter_1", "adapter_1": "adapter_0"}.get(name)
        for n, flag in ((name, True), (other, False)):
            if n is None:
                continue
            for part in ("down", "up"):
                for p in getattr(self, f"{n}_{part}").parameters():
                    p.requires_grad = flag

    def activate_gating(self):
        self.gating = True

    def deactivate_gating(self):
        self.gating = False

    def _branch(self, name, h):
        down = F.relu(getattr(self, f"{name}_down")(h))              # adapter.py:127-128 / 137-138
        return getattr(self, f"{name}_up")(down)                     # adapter.py:129 / 140

    def forward(self, hidden_states, input_tensor):
        if not self.gating:
            return input_tensor + self._branch(self._active, hidden_states)       # adapter.py:131
        second = "adapter_2" if hasattr(self, "adapter_2_down") else "adapter_1"
        agg = 0.5 * self._branch("adapter_0", hidden_states)                        # adapter.py:144, 118-122
        agg = agg + 0.5 * self._branch(second, hidden_states)
        return input_tensor + agg * self.scaling                                    # adapter.py:146


class OracleAdapteredViltOutput(nn.Module):
    """adaptered_output.py:67-79."""

    def __init__(self, layer, rank):
        super().__init__()
        self.layer = layer
        self.adapter = OracleAdapter(["adapter_0", "adapter_1", "adapter_2"], rank)

    def forward(self, hidden_states, input_tensor):
        hidden_states = self.layer.dense(hidden_states)
        hidden_states = self.layer.dropout(hidden_states)
        hidden_states = hidden_states + input_tensor
        return self.adapter(hidden_states, hidden_states)


class OracleLearner(nn.Module):
    """Tensor-input stand-in for ViltContinualLearner (vilt.py:152-382) with identical state-dict keys."""

    def __init__(self, rank, tasks=("art",), num_labels=100):
        super().__init__()
        from transformers import ViltConfig, ViltModel
        enc = nn.Module()
        enc.vilt = ViltModel(ViltConfig())
        emb = enc.vilt.embeddings.token_type_embeddings.weight.data              # vilt.py:102-113
        enc.vilt.embeddings.token_type_embeddings = nn.Embedding(3, 768)
        enc.vilt.embeddings.token_type_embeddings.weight.data[:2] = emb[:2]
        enc.vilt.embeddings.token_type_embeddings.weight.data[2] = emb[1]
        self.vilt_encoder = enc
        for i in range(12):                                                       # vilt.py:356-361
            self.vilt_encoder.vilt.encoder.layer[i].output = OracleAdapteredViltOutput(
                self.vilt_encoder.vilt.encoder.layer[i].output, rank)
        self.task_layer = nn.ModuleDict({t: nn.Sequential(OrderedDict([          # vilt.py:202-209
            ("clf_fc0", nn.Linear(768, 1536)), ("clf_norm0", nn.LayerNorm(1536)),
            ("clf_actv0", nn.GELU()), ("clf_fc1", nn.Linear(1536, num_labels))])) for t in tasks})

    def _adapters(self):
        return [l.output.adapter for l in self.vilt_encoder.vilt.encoder.layer]

    def set_active_adapter(self, name):
        for a in self._adapters():
            a.set_active_adapter(name)

    def activate_gating(self):
        for a in self._adapters():
            a.activate_gating()

    def deactivate_gating(self):
        for a in self._adapters():
            a.deactivate_gating()

    def forward(self, task_key, **enc):
        enc.pop("dense_masks", None)
        pooled = self.vilt_encoder.vilt(**enc).pooler_output
        return pooled, self.task_layer[task_key](pooled)

    def prepare_dat(self):
        """main.py:138-139,157-159,248-250 + task_trainer.py:36-45."""
        for p in self.parameters():
            p.requires_grad = False
        for n, p in self.named_parameters():
            if "adapter" in n or "task" in n:
                p.requires_grad = True
        sd = self.state_dict()
        for name in sd:
            if "adapter_1" in name:
                sd[name.replace("adapter_1", "adapter_2")].data.copy_(sd[name].data)
        for n, p in self.named_parameters():
            if "adapter_2" in n:
                p.requires_grad = False
        return self


def kl_loss(output, target, temp=3):
    """task_trainer.py:506-516."""
    dim = -1 if output.shape[-1] > 3000 else 1
    p = F.log_softmax(output / temp, dim=dim)
    q = F.softmax(target / temp, dim=dim)
    return F.kl_div(p, q, reduction="batchmean") * temp ** 2


def create_optimizer(model, lr, weight_decay=1e-2, eps=1e-8):
    """task_trainer.py:477-504."""
    no_decay = ["bias", "LayerNorm.weight"]
    groups = [
        {"params": [p for n, p in model.named_parameters()
                    if (not any(nd in n for nd in no_decay)) and p.requires_grad], "weight_decay": weight_decay},
        {"params": [p for n, p in model.named_parameters()
                    if any(nd in n for nd in no_decay) and p.requires_grad], "weight_decay": 0.0},
    ]
    return AdamW(groups, lr=lr, eps=eps, betas=(0.9, 0.98))


def create_scheduler(optimizer, max_steps, warmup_ratio=0.1):
    """task_trainer.py:53-59."""
    from transformers import get_polynomial_decay_schedule_with_warmup
    return get_polynomial_decay_schedule_with_warmup(optimizer, num_warmup_steps=int(max_steps * warmup_ratio),
                                                     num_training_steps=max_steps, lr_end=0, power=1)


def train_step(model, task_key, batch, optimizer, scheduler, temp=3):
    """task_trainer.py:280-330 (dat branch, ViLT).  Returns (loss_0, (logits_all, logits_1, logits_0))."""
    enc = dict(batch["encodings"])
    target = batch["target_scores"]
    crit = nn.BCEWithLogitsLoss(reduction="mean")                     # train_vqa_crossvqa.py:237

    with torch.no_grad():                                             # :283-287
        model.activate_gating()
        _, logits_all = model(task_key, **enc)

    model.deactivate_gating()                                         # :290-291
    model.set_active_adapter("adapter_1")
    _, logits_1 = model(task_key, **enc)
    loss_1 = crit(logits_1, target) * target.shape[1]                 # :299
    L_1 = (loss_1 + kl_loss(logits_1, logits_all.clone().detach(), temp)) / 2   # :300-301
    L_1.backward()
    optimizer.step()
    scheduler.step()
    optimizer.zero_grad()

    model.activate_gating()                                           # :311-312
    model.set_active_adapter("adapter_0")
    _, logits_0 = model(task_key, **enc)
    loss_0 = crit(logits_0, target) * target.shape[1]                 # :319
    L_0 = (loss_0 + kl_loss(logits_0, logits_1.clone().detach(), temp)) / 2     # :320-321
    L_0.backward()
    optimizer.step()
    scheduler.step()
    optimizer.zero_grad()
    return loss_0.detach(), (logits_all.detach(), logits_1.detach(), logits_0.detach())
