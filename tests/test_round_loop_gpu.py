"""Round-level parity (level 4): the executed round loop ``feddat_b200.train.main.main``
for 2 rounds x 3 clients with the eval after round 0 (reference src/train/main.py:453-558,
task_trainer.py:211-246).  Pins: (i) the FedAvg result of every round is bit-identical to the oracle's
``get_average_net`` over the clients' adapter_1 snapshots; (ii) every client's optimizer holds adapter_1
(the resident model must not leak the previous client's requires_grad flags); (iii) SURVEY.md F8 -- after
the round-0 eval adapter_0 is frozen on the server model, so round-1 optimizers lack it; (iv) personal
parameters stay per client.  GPU only (the DAT operator has no CPU path)."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

ARGV = ["--encoder_name", "vilt", "--pretrained_model_name", "random", "--climb_data_dir", "synthetic",
        "--do_train", "--output_dir", "/tmp/feddat_round_test", "--optimizer_mode", "dat",
        "--ordered_cl_tasks", "art,abstract,vizwiz", "--comm_round", "2", "--local_epochs", "1",
        "--batch_size", "2", "--val_batch_size", "2", "--synthetic_batches", "2", "--image_size", "224",
        "--text_len", "16", "--adapter_rank", "32", "--adapter_config", "pfeiffer", "--lr", "1e-3",
        "--seed", "5", "--num_epochs", "2"]


@pytest.mark.parametrize("extra", [[], ["--cuda_graph"]], ids=["eager", "graphed"])
def test_two_rounds_three_clients(extra):
    from feddat_b200.train.main import main
    rec = {}
    assert main(ARGV + extra, record=rec) == 0
    clients = ["art", "abstract", "vizwiz"]
    for rnd in (0, 1):
        flats = [f.numpy() for f in rec["client_flats"][rnd]]
        assert len(flats) == 3
        want = oracle.get_average_net(flats, [1, 1, 1])
        got = rec["global_flat"][rnd].numpy()
        assert np.array_equal(got, want), f"round {rnd}: FedAvg differs from the reference expression"
        # the clients really trained adapter_1 (different data per client -> different snapshots)
        assert not np.array_equal(flats[0], flats[1]) and not np.array_equal(flats[1], flats[2])
    for c in clients:
        names0 = rec["optimizer_names"][(0, c)]
        names1 = rec["optimizer_names"][(1, c)]
        assert any("adapter_1" in n for n in names0) and any("adapter_1" in n for n in names1), c
        assert any("adapter_0" in n for n in names0), c                 # round 0: server flags as prepared
        assert not any("adapter_0" in n for n in names1), c             # F8: frozen by the round-0 eval
        assert not any("adapter_2" in n for n in names0 + names1), c
        # 12 sites x 2 adapters x 4 tensors + the 6-tensor head of THIS client... the reference enables every
        # client's head ('task' substring, main.py:248-250): 3 heads x 6 tensors
        assert len(names0) == 12 * 2 * 4 + 3 * 6 and len(names1) == 12 * 1 * 4 + 3 * 6
    assert set(rec["eval_scores"]) == {0, 1}                             # round 0 and the last round
    assert all(np.isfinite(v) for s in rec["eval_scores"].values() for v in s)


def test_fix_flag_restores_adapter0():
    from feddat_b200.train.main import main
    rec = {}
    assert main(ARGV + ["--fix_adapter0_optimizer"], record=rec) == 0
    assert any("adapter_0" in n for n in rec["optimizer_names"][(1, "art")])
