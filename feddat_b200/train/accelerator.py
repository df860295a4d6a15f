"""Minimal stand-in for ``accelerate.Accelerator`` (absent from this image) exposing
exactly the surface the reference's round loop and trainer touch: ``device, process_index,
is_main_process, prepare, backward, gather, wait_for_everyone, unwrap_model, free_memory, log,
init_trackers``.  One process per GPU; no DDP inside a client (reference default num_processes: 1).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist
import torch.nn as nn


class ModuleWrapper(nn.Module):
    """Plays the role of the DDP wrapper: the reference calls ``model.module.<hook>()``
    (task_trainer.py:283-312) and ``model(...)``."""

    def __init__(self, module: nn.Module):
        super().__init__()
        self.module = module

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)


class Accelerator:
    def __init__(self, device=None, **_ignored):
        self.process_index = int(os.environ.get("RANK", "0"))
        self.local_process_index = int(os.environ.get("LOCAL_RANK", "0"))
        self.num_processes = int(os.environ.get("WORLD_SIZE", "1"))
        if device is None:
            device = torch.device("cuda", self.local_process_index) if torch.cuda.is_available() else torch.device("cpu")
        self.device = torch.device(device)
        if self.device.type == "cuda":
            if self.device.index is None:
                self.device = torch.device("cuda", torch.cuda.current_device())
            torch.cuda.set_device(self.device)

    @property
    def is_main_process(self):
        return self.process_index == 0

    def prepare(self, *objs):
        out = []
        for o in objs:
            if isinstance(o, nn.Module) and not isinstance(o, ModuleWrapper):
                out.append(ModuleWrapper(o.to(self.device)))
            else:
                out.append(o)
        return out[0] if len(out) == 1 else tuple(out)

    def unwrap_model(self, model):
        return model.module if isinstance(model, ModuleWrapper) else model

    def backward(self, loss):
        loss.backward()

    def gather(self, t):
        if self.num_processes == 1 or not dist.is_initialized():
            return t
        outs = [torch.empty_like(t) for _ in range(self.num_processes)]
        dist.all_gather(outs, t.contiguous())
        return torch.cat(outs, dim=0)

    def wait_for_everyone(self):
        """accelerate's barrier over the processes that run ONE client's training (reference task_trainer.py:107, called
        from inside ``TaskTrainer.train``).  Here a client lives on exactly one process, so there
        is nobody to wait for -- and it must NOT be a barrier over all ranks: ranks hold different numbers of clients
        (5 domain clients on 2 GPUs = 3 + 2), so a global barrier per client leaves the ranks one collective apart and
        the job hangs until the NCCL watchdog aborts it."""

    def barrier_all_ranks(self):
        """The round boundary: every rank has trained its clients (feddat_b200/train/main.py)."""
        if self.num_processes > 1 and dist.is_initialized():
            dist.barrier()

    def free_memory(self):
        pass

    def log(self, *_a, **_k):
        pass

    def init_trackers(self, *_a, **_k):
        pass
