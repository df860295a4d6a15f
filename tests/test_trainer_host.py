"""Host logic of TaskTrainer.train_step (the MKD schedule of reference task_trainer.py:280-330) on
CPU with a stand-in model; the fused loss kernel is replaced by the oracle for this test only."""
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn

import oracle
from feddat_b200.train import task_trainer as tt
from feddat_b200.train.accelerator import Accelerator


def _oracle_mkd(logits, teacher, target, temp, kl_weight=0.5, task_weight=0.5, need_grad=True):
    lg, te = logits.detach().numpy(), teacher.detach().numpy()
    kl, gkl = oracle.kl_loss(lg, te, temp, with_grad=True)
    task, gtask = (0.0, 0.0) if target is None else oracle.bce_with_logits_times_c(lg, target.numpy(), with_grad=True)
    loss3 = torch.tensor([kl_weight * kl + task_weight * task, kl, task], dtype=torch.float32)
    return loss3, torch.from_numpy(np.asarray(kl_weight * gkl + task_weight * gtask, np.float32))


class FakeLearner(nn.Module):
    """Two 'adapters' + head; records the hook calls the trainer makes."""

    def __init__(self):
        super().__init__()
        self.adapter_0 = nn.Linear(6, 6)
        self.adapter_1 = nn.Linear(6, 6)
        self.adapter_2 = nn.Linear(6, 6)
        self.task_head = nn.Linear(6, 5)
        self.gating, self.active, self.log = False, None, []

    def activate_gating(self):
        self.gating = True; self.log.append("gate_on")

    def deactivate_gating(self):
        self.gating = False; self.log.append("gate_off")

    def set_active_adapter(self, name):
        self.active = name; self.log.append(name)
        for p in self.adapter_0.parameters():
            p.requires_grad = name == "adapter_0"
        for p in self.adapter_1.parameters():
            p.requires_grad = name == "adapter_1"

    def forward(self, task_key, x):
        self.log.append("fwd")
        if self.gating:
            h = x + 0.5 * self.adapter_0(x) + 0.5 * self.adapter_2(x)
        else:
            h = x + getattr(self, self.active)(x)
        return h, self.task_head(h)


def test_dat_schedule_and_losses(monkeypatch):
    monkeypatch.setattr(tt.ops, "mkd_loss", _oracle_mkd)
    torch.manual_seed(0)
    m = FakeLearner()
    for p in m.adapter_2.parameters():
        p.requires_grad = False
    tr = tt.TaskTrainer()
    tr.args = SimpleNamespace(optimizer_mode="dat", encoder_name="vilt")
    tr.accelerator = Accelerator(device="cpu")
    tr.device, tr.task_key = torch.device("cpu"), "t"
    tr.batch2inputs_converter = lambda b: {"x": b["x"]}
    tr.loss_criterion = nn.BCEWithLogitsLoss(reduction="mean")
    tr.weight_decay, tr.lr, tr.adam_epsilon, tr.kl_temp = 1e-2, 1e-2, 1e-8, 3
    wrapped = tr.accelerator.prepare(m)
    opt = tr.create_optimizer(wrapped)
    steps = []
    orig = opt.step
    opt.step = lambda *a, **k: (steps.append(len(m.log)), orig(*a, **k))[1]
    x = torch.randn(4, 6)
    target = (torch.rand(4, 5) < 0.3).float() * 0.6
    w0 = {n: p.detach().clone() for n, p in m.named_parameters()}
    loss_0 = tr.train_step(wrapped, 0, {"x": x, "target_scores": target}, opt, None)
    # hook order of task_trainer.py:283-315
    assert m.log == ["gate_on", "fwd", "gate_off", "adapter_1", "fwd", "gate_on", "adapter_0", "fwd"]
    assert len(steps) == 2                                     # optimizer stepped once per pass B and C
    moved = {n: not torch.equal(p, w0[n]) for n, p in m.named_parameters()}
    assert moved["adapter_1.weight"] and moved["adapter_0.weight"] and moved["task_head.weight"]
    assert not moved["adapter_2.weight"]
    # loss_0 is the TASK term of pass C (reference returns loss_0, not L_0)
    logits_0 = tr.last_logits[2].detach().numpy()
    assert abs(loss_0.item() - oracle.bce_with_logits_times_c(logits_0, target.numpy())) < 1e-4
    # pass C distils from pass B's logits: L_0 = (task + kl(logits_0, logits_1)) / 2
    want_L0 = oracle.mkd_total(logits_0, tr.last_logits[1].detach().numpy(), target.numpy(), 3)[0]
    assert abs(tr.last_objectives[1].item() - want_L0) < 1e-4
    assert all(p.grad is None for p in m.parameters())         # zero_grad -> None (torch >= 2 contract)


class ReusableLearner(FakeLearner):
    """FakeLearner with the encode / classify split the ViLT learner exposes (no dropout)."""

    def encode(self, x):
        self.log.append("enc")
        assert self.gating
        return x + 0.5 * self.adapter_0(x) + 0.5 * self.adapter_2(x)

    def classify(self, task_key, h):
        return self.task_head(h)

    def gating_forward_is_reusable(self):
        return True


def _one_step(learner_cls, reuse):
    torch.manual_seed(0)
    m = learner_cls()
    for p in m.adapter_2.parameters():
        p.requires_grad = False
    tr = tt.TaskTrainer()
    tr.reuse_gating_forward = reuse
    tr.args = SimpleNamespace(optimizer_mode="dat", encoder_name="vilt")
    tr.accelerator = Accelerator(device="cpu")
    tr.device, tr.task_key = torch.device("cpu"), "t"
    tr.batch2inputs_converter = lambda b: {"x": b["x"]}
    tr.loss_criterion = nn.BCEWithLogitsLoss(reduction="mean")
    tr.weight_decay, tr.lr, tr.adam_epsilon, tr.kl_temp = 1e-2, 1e-2, 1e-8, 3
    wrapped = tr.accelerator.prepare(m)
    opt = tr.create_optimizer(wrapped)
    g = torch.Generator().manual_seed(1)
    losses = []
    for i in range(3):
        x = torch.randn(4, 6, generator=g)
        target = (torch.rand(4, 5, generator=g) < 0.3).float() * 0.6
        losses.append(tr.train_step(wrapped, i, {"x": x, "target_scores": target}, opt, None).item())
    return m, losses, tr


def test_shared_gating_forward_equals_three_forward_schedule(monkeypatch):
    """Passes A and C share one encoder forward -- same losses, same parameters, same
    final requires_grad state as the reference's three-forward schedule."""
    monkeypatch.setattr(tt.ops, "mkd_loss", _oracle_mkd)
    m_ref, l_ref, _ = _one_step(ReusableLearner, reuse=False)
    m_new, l_new, _ = _one_step(ReusableLearner, reuse=True)
    assert m_ref.log.count("fwd") == 9 and m_new.log.count("fwd") == 3 and m_new.log.count("enc") == 3
    np.testing.assert_allclose(l_new, l_ref, rtol=1e-6)
    for (n, p), (_, q) in zip(m_ref.named_parameters(), m_new.named_parameters()):
        assert torch.allclose(p, q, rtol=1e-6, atol=1e-7), n
        assert p.requires_grad == q.requires_grad, n


def test_split_step_equals_the_reference_two_optimizer_steps():
    """The batched MKD schedule's optimizer choreography (head stepped early, then split_step) leaves every
    parameter exactly where the reference's "step B, scheduler, step C, scheduler" leaves it (AdamW is
    per-parameter; weight decay and bias correction included)."""
    import copy
    from torch.optim.lr_scheduler import LambdaLR
    from feddat_b200.train.task_trainer import split_step
    torch.manual_seed(0)
    base = {"a1": torch.randn(4, 3), "a0": torch.randn(4, 3), "head": torch.randn(5)}

    def make():
        ps = {k: torch.nn.Parameter(v.clone()) for k, v in base.items()}
        opt = torch.optim.AdamW([{"params": [ps["a1"], ps["a0"]], "weight_decay": 1e-2},
                                 {"params": [ps["head"]], "weight_decay": 0.0}], lr=1e-2, betas=(0.9, 0.98))
        return ps, opt, LambdaLR(opt, lambda e: 1.0 / (1 + e))

    grads = [{k: torch.randn_like(v) for k, v in base.items()} | {"head2": torch.randn(5)} for _ in range(3)]
    ref, o1, s1 = make()
    for g in grads:                                        # reference order (task_trainer.py:290-328)
        ref["a1"].grad, ref["head"].grad = g["a1"].clone(), g["head"].clone()
        o1.step(); s1.step(); o1.zero_grad()
        ref["a0"].grad, ref["head"].grad = g["a0"].clone(), g["head2"].clone()
        o1.step(); s1.step(); o1.zero_grad()
    new, o2, s2 = make()
    for g in grads:                                        # batched order
        new["head"].grad = g["head"].clone()
        o2.step(); o2.zero_grad()                          # head early, first learning rate
        new["a1"].grad, new["a0"].grad, new["head"].grad = g["a1"].clone(), g["a0"].clone(), g["head2"].clone()
        split_step(o2, s2, [new["a1"]])
    for k in base:
        assert torch.equal(ref[k], new[k]), k
    assert s1.last_epoch == s2.last_epoch == 6
