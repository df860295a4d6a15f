"""Host logic of the round boundary: flat communicated buffer + FedAvg, single process and
world_size-2 (gloo, CPU).  The fedavg CUDA kernel cannot run here, so ``ops.fedavg`` is replaced by
the oracle for these tests only -- what is under test is the partition / partial-sum / allreduce
plumbing around it."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

import oracle
from feddat_b200.train import fedavg as fa


def _oracle_fedavg(bufs, nums, out, total=0.0):
    total = float(total) if total and total > 0 else float(sum(nums))
    acc = np.zeros(out.numel(), np.float32)
    for b, n in zip(bufs, nums):
        acc = acc + (b.detach().cpu().numpy().astype(np.float32) * np.float32(n)) / np.float32(total)
    out.copy_(torch.from_numpy(acc))
    return out


class Tiny(nn.Module):
    def __init__(self):
        super().__init__()
        self.adapter_0_down = nn.Linear(8, 3)
        self.adapter_1_down = nn.Linear(8, 3)
        self.adapter_1_up = nn.Linear(3, 8)
        self.clf_adapter_1 = nn.Linear(2, 2)          # 'clf' keys are skipped (main.py:54)


def test_flat_buffer_aliases_parameters():
    m = Tiny()
    names = fa.comm_state_dict_names(m)
    assert all("adapter_1" in n for n in names) and len(names) == 6
    before = {n: p.detach().clone() for n, p in m.named_parameters()}
    flat = fa.FlatCommBuffer(m, names)
    assert flat.numel == 3 * 8 + 3 + 8 * 3 + 8 and flat.flat.numel() % 4 == 0
    assert all("clf" not in n for n in flat.names)
    for n, p in m.named_parameters():
        assert torch.equal(p, before[n])               # values preserved
    flat.flat.mul_(2.0)                                # writing the flat buffer moves the parameters
    assert torch.equal(m.adapter_1_down.weight, before["adapter_1_down.weight"] * 2)
    assert torch.equal(m.adapter_0_down.weight, before["adapter_0_down.weight"])
    with torch.no_grad():
        m.adapter_1_up.bias.add_(1.0)                  # and optimizer-style in-place updates land in it
    sl = flat.slices["adapter_1_up.bias"]
    assert torch.equal(flat.flat[sl], m.adapter_1_up.bias.detach())


def test_single_process_average_matches_reference_order(monkeypatch):
    monkeypatch.setattr(fa.ops, "fedavg", _oracle_fedavg)
    torch.manual_seed(0)
    server = Tiny()
    flat = fa.FlatCommBuffer(server, fa.comm_state_dict_names(server))
    clients = [torch.randn_like(flat.flat) for _ in range(3)]
    fa.get_average_net_flat(flat, clients, [3, 1, 2])
    want = oracle.get_average_net([c.numpy() for c in clients], [3, 1, 2])
    assert np.array_equal(flat.flat.numpy(), want)


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fa.ops.fedavg = _oracle_fedavg
    torch.manual_seed(7)                               # same "server" everywhere
    server = Tiny()
    flat = fa.FlatCommBuffer(server, fa.comm_state_dict_names(server))
    n_clients = 5                                      # clients > ranks: rank r trains c with c % world == r
    g = torch.Generator().manual_seed(100)
    all_clients = [torch.randn(flat.flat.numel(), generator=g) for _ in range(n_clients)]
    mine = [all_clients[c] for c in range(n_clients) if c % world == rank]
    fa.get_average_net_flat(flat, mine, [1.0] * len(mine), total=float(n_clients))
    ret[rank] = flat.flat.clone().numpy()
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_allreduce_average_matches_single_process():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    g = torch.Generator().manual_seed(100)
    n = ret[0].size
    all_clients = [torch.randn(n, generator=g).numpy() for _ in range(5)]
    want = oracle.get_average_net(all_clients, [1] * 5)
    assert np.array_equal(ret[0], ret[1])                       # every rank holds the same global adapter
    np.testing.assert_allclose(ret[0], want, rtol=1e-6, atol=1e-7)   # summation order differs across ranks


def _uneven_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from feddat_b200.train.accelerator import Accelerator
    acc = Accelerator(device="cpu")
    server = Tiny()
    flat = fa.FlatCommBuffer(server, fa.comm_state_dict_names(server))
    fa.ops.fedavg = _oracle_fedavg
    n_clients = 5
    for _round in range(2):
        mine = [c for c in range(n_clients) if c % world == rank]             # 3 clients on rank 0, 2 on rank 1
        for _c in mine:
            acc.wait_for_everyone()       # what TaskTrainer.train calls once per CLIENT (task_trainer.py:107)
        fa.get_average_net_flat(flat, [torch.full_like(flat.flat, float(c)) for c in mine], [1.0] * len(mine),
                                total=float(n_clients))
        acc.barrier_all_ranks()
    ret[rank] = float(flat.flat[0])
    dist.destroy_process_group()


def test_uneven_client_counts_do_not_desynchronise_the_ranks():
    """5 clients on 2 ranks (the default ``src/train_vilt.sh`` task list on 2 GPUs): the per-client
    ``wait_for_everyone`` inside ``TaskTrainer.train`` is local to the client's process, so the ranks meet only at the
    round boundary.  (As a global barrier it left the ranks one collective apart: a hang until the NCCL watchdog.)"""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_uneven_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] == ret[1] == 2.0                       # mean of clients 0 .. 4
