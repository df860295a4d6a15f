"""Host logic of the federated round loop (reference src/train/main.py:453-558) with the resident model
that replaces ``copy.deepcopy(model)`` per client: every client must start from the SERVER model's
``requires_grad`` flags (what the deepcopy carries), not from the flags the previous client's last
``train_step`` left behind."""
from types import SimpleNamespace

import torch

from feddat_b200.train.accelerator import Accelerator
from feddat_b200.train.main import grad_flags, restore_grad_flags
from feddat_b200.train.prepare import default_args, prepare_model
from feddat_b200.train.task_trainer import TaskTrainer


def _trainer():
    tr = TaskTrainer()
    tr.args = SimpleNamespace(optimizer_mode="dat", encoder_name="vilt", debug=0)
    tr.accelerator = Accelerator(device="cpu")
    tr.device = torch.device("cpu")
    tr.weight_decay, tr.lr, tr.adam_epsilon = 1e-2, 1e-4, 1e-8
    return tr


def _optimizer_names(tr, model):
    opt = tr.create_optimizer(tr.accelerator.prepare(model))
    ids = {id(p) for g in opt.param_groups for p in g["params"]}
    return [n for n, p in model.named_parameters() if id(p) in ids]


def _end_of_train_step(model):
    """The mode switches every dat train_step ends with (task_trainer.py:311-312)."""
    model.activate_gating()
    model.set_active_adapter("adapter_0")


def _prologue(model):
    """TaskTrainer.train :43-45."""
    for n, p in model.named_parameters():
        if "adapter_2" in n:
            p.requires_grad = False


def test_every_client_optimizer_holds_adapter_1():
    torch.manual_seed(0)
    model = prepare_model(default_args(ordered_cl_tasks=["art", "abstract"], adapter_rank=16), place=False)
    server = grad_flags(model)
    tr = _trainer()
    seen = []
    for _client in range(3):
        restore_grad_flags(model, server)            # what main() does where the reference deep-copies
        _prologue(model)
        names = _optimizer_names(tr, model)
        seen.append(names)
        _end_of_train_step(model)                    # leaves adapter_1.requires_grad = False behind
    for names in seen:
        assert any("adapter_1" in n for n in names)
        assert any("adapter_0" in n for n in names)
        assert not any("adapter_2" in n for n in names)
    assert seen[0] == seen[1] == seen[2]
    # without the restore the second client loses adapter_1 (the round-1 bug this guards against)
    _prologue(model)
    assert not any("adapter_1" in n for n in _optimizer_names(tr, model))


def test_eval_side_effect_is_what_later_rounds_inherit():
    """eval() on the server model ends on set_active_adapter('adapter_1')
    (task_trainer.py:236-243), so every later deepcopy has adapter_0 frozen and its optimizer lacks it."""
    torch.manual_seed(0)
    model = prepare_model(default_args(ordered_cl_tasks=["art"], adapter_rank=16), place=False)
    tr = _trainer()
    # eval's hook sequence
    model.activate_gating(); model.deactivate_gating(); model.set_active_adapter("adapter_0")
    model.deactivate_gating(); model.set_active_adapter("adapter_1")
    server = grad_flags(model)
    for _client in range(2):
        restore_grad_flags(model, server)
        _prologue(model)
        names = _optimizer_names(tr, model)
        assert any("adapter_1" in n for n in names) and not any("adapter_0" in n for n in names)
        _end_of_train_step(model)
