"""Round-boundary FedAvg (mirror of reference src/train/main.py:50-65, key selection :154-163).

B200-native layout: every communicated parameter (state-dict keys containing 'adapter_1', skipping
keys with 'clf' as main.py:54 does) is re-homed as a VIEW into one contiguous fp32 buffer per model
(``FlatCommBuffer``).  A client's contribution is then one device-to-device snapshot of that buffer,
the average of the clients sharing a GPU is ONE fedavg kernel launch, and the exchange across GPUs
is ONE NCCL allreduce(SUM) of the pre-weighted partial sums over NVLink -- the only collective on
the path.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.distributed as dist
import torch.nn as nn

from .. import ops


def comm_state_dict_names(model: nn.Module, shared_params_names=("adapter_1",)) -> List[str]:
    """main.py:160-163."""
    return [n for n in model.state_dict().keys() if any(sn in n for sn in shared_params_names)]


class FlatCommBuffer:
    """Owns the flat fp32 buffer behind the communicated parameters of ``model``."""

    def __init__(self, model: nn.Module, names: Sequence[str]):
        self.names = [n for n in names if "clf" not in n]            # main.py:54
        params = dict(model.named_parameters())
        sd = model.state_dict()
        sizes = [sd[n].numel() for n in self.names]
        total = sum(sizes)
        pad = (-total) % 4                                           # float4 kernel / NCCL alignment
        ref = sd[self.names[0]]
        if ref.dtype != torch.float32:
            raise TypeError("communicated adapter parameters must be fp32 masters")
        self.flat = torch.zeros(total + pad, device=ref.device, dtype=torch.float32)
        self.numel = total
        self.slices: Dict[str, slice] = {}
        off = 0
        for n, sz in zip(self.names, sizes):
            view = self.flat[off: off + sz].view(sd[n].shape)
            view.copy_(sd[n])
            if n in params:
                params[n].data = view                                # parameter now aliases the flat buffer
            else:
                raise KeyError(f"{n} is not a parameter")
            self.slices[n] = slice(off, off + sz)
            off += sz

    def snapshot(self) -> torch.Tensor:
        return self.flat.clone()

    def load(self, flat: torch.Tensor) -> None:
        self.flat.copy_(flat)


def get_average_net_flat(server: FlatCommBuffer, client_flats: Sequence[torch.Tensor], nums: Sequence[float],
                         total: float = 0.0, group=None) -> None:
    """server <- sum_c client_c * num_c / total.  Single process: bit-identical to the reference
    expression (same operation order).  With torch.distributed initialised, ``client_flats`` are this
    rank's clients, ``total`` the global sum of nums, and one allreduce(SUM) finishes the average."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        ops.fedavg(list(client_flats), list(nums), server.flat, total)
        return
    if total <= 0:
        raise ValueError("multi-rank FedAvg needs the global total weight")
    if len(client_flats) > 0:
        ops.fedavg(list(client_flats), list(nums), server.flat, total)
    else:
        server.flat.zero_()
    dist.all_reduce(server.flat, op=dist.ReduceOp.SUM, group=group)


def get_average_net(server: nn.Module, c_models: Sequence[Dict[str, torch.Tensor]], nums, ordered_tasks=None,
                    device=None):
    """Signature-compatible with the reference (main.py:50): ``c_models`` are dicts key -> tensor.
    Keys are averaged through the same fedavg kernel, one launch per key."""
    total_names = getattr(server, "comm_state_dict_names", None) or comm_state_dict_names(server)
    sd = server.state_dict()
    with torch.no_grad():
        for key in total_names:
            if "clf" in key:
                continue
            if "num_batches_tracked" in key:
                sd[key].data.copy_(c_models[0][key])
                continue
            srcs = [m[key].contiguous().float().view(-1) for m in c_models]
            out = torch.empty_like(srcs[0])
            n = out.numel()
            if n % 4 or any(s.data_ptr() % 16 for s in srcs):
                pad = (-n) % 4
                srcs = [torch.cat([s, s.new_zeros(pad)]) for s in srcs]
                out = torch.empty_like(srcs[0])
            ops.fedavg(srcs, list(nums), out)
            sd[key].data.copy_(out[:n].view(sd[key].shape))
    return server
