"""Pins oracle/step_oracle.py (the CPU port used as cpu_baseline / --impl reference) against the
golden produced by the reference's own TaskTrainer.train_step + Adapter.  CPU only (~20 s)."""
from pathlib import Path

import numpy as np
import torch

from oracle import step_oracle


def test_step_oracle_matches_reference_trainer():
    from feddat_b200.synthetic import make_vilt_batch
    from feddat_b200.train.prepare import default_args, prepare_model
    gold = np.load(Path(__file__).resolve().parent / "golden" / "step_golden.npz")
    seed, rank, steps, max_steps, B, T, H, C = (int(v) for v in gold["meta"])
    torch.manual_seed(seed)
    ours = prepare_model(default_args(ordered_cl_tasks=["art"], adapter_rank=rank), place=False)
    model = step_oracle.OracleLearner(rank)
    missing, unexpected = model.load_state_dict(ours.state_dict(), strict=False)
    assert not unexpected and all("_ids" in k for k in missing)
    model.prepare_dat()
    opt = step_oracle.create_optimizer(model, float(gold["lr"]))
    assert sum(len(g["params"]) for g in opt.param_groups) == int(gold["n_optimizer_tensors"])
    sched = step_oracle.create_scheduler(opt, max_steps)
    model.train()
    for step in range(steps):
        batch = make_vilt_batch(B, T, H, C, seed=seed + step)
        loss_0, logits = step_oracle.train_step(model, "art", batch, opt, sched)
        assert abs(loss_0.item() - float(gold[f"step{step}/loss_0"])) < 1e-3 * float(gold[f"step{step}/loss_0"])
        for name, t in zip(("logits_all", "logits_1", "logits_0"), logits):
            np.testing.assert_allclose(t.numpy(), gold[f"step{step}/{name}"], rtol=2e-3, atol=2e-4)
