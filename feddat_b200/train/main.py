"""Federated round loop (mirror of reference src/train/main.py: argparse :262-323, personal-parameter
stash :440-450, round loop :453-518, periodic eval :520-558), B200-native placement:

  * one process per GPU (torchrun), client c trains on rank c % world; no DDP inside a client
    (the reference ships num_processes: 1 and simulates clients sequentially in one process)
  * ONE resident model per GPU: instead of ``copy.deepcopy(model)`` per client per round
    (main.py:472) the client's personal parameters (keys containing 'task' | 'adapter_0' |
    'adapter_2') are copied in and out of the resident model in place, and the shared 'adapter_1'
    parameters are views into one flat fp32 buffer
  * the round boundary is one fedavg kernel launch over this rank's clients + ONE NCCL allreduce of
    the flat buffer (get_average_net, main.py:50-65, equal weights nums = [1, ...] main.py:455)

Flags are the reference's (train_vilt.sh:1-19 parses unchanged, including the prefix abbreviation
``--comm_round``), plus --adapter_rank / --adapter_activation / --kl_temp and the synthetic-data
knobs (there are no datasets on the box: ``--climb_data_dir`` is accepted and ignored).
"""
from __future__ import annotations

import argparse
import logging
import os
import sys
import time

import torch
import torch.distributed as dist

from ..configs.adapter_configs import ADAPTER_MAP
from ..configs.model_configs import ALLOWED_CL_ENCODERS, model_configs
from ..configs.task_configs_fed import DOMAIN_TASKS, task_configs
from .accelerator import Accelerator
from .fedavg import FlatCommBuffer, get_average_net_flat
from .prepare import prepare_model
from .train_vqa_synthetic import VQATrainerSynthetic


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("--encoder_name", default=None, type=str, required=True, choices=ALLOWED_CL_ENCODERS)
    p.add_argument("--portion", default=1.0, type=float)
    p.add_argument("--optimizer_mode", default="none", type=str)
    p.add_argument("--pretrained_model_name", default=None, type=str, required=True)
    p.add_argument("--climb_data_dir", type=str, required=True, default="")
    p.add_argument("--debug", type=int, default=0)
    p.add_argument("--do_single", action="store_true")
    p.add_argument("--do_train", action="store_true")
    p.add_argument("--do_eval", action="store_true")
    p.add_argument("--do_test", action="store_true")
    p.add_argument("--adapter_config", choices=list(ADAPTER_MAP.keys()))
    p.add_argument("--adapter_reduction_factor", type=int, default=0)
    p.add_argument("--layers_to_freeze", type=int, default=0)
    p.add_argument("--output_dir", type=str, required=True)
    p.add_argument("--do_wandb_logging", action="store_true")
    p.add_argument("--wandb_freq", type=int, default=100)
    p.add_argument("--comm_rounds", type=int, default=20)
    p.add_argument("--local_epochs", type=int, default=1)
    p.add_argument("--batch_size", type=int, default=32)
    p.add_argument("--num_epochs", type=int, default=15)
    p.add_argument("--val_batch_size", type=int, default=1)
    p.add_argument("--num_workers", type=int, default=2)
    p.add_argument("--seed", type=int, default=42)
    p.add_argument("--ordered_cl_tasks", type=str)
    p.add_argument("--lr", default=None, type=float)
    p.add_argument("--splits", nargs="*", default=["train", "val"])
    p.add_argument("--checkpoint", type=str, default=None)
    p.add_argument("--model_path", type=str, default=None)
    # new (B200 build)
    p.add_argument("--adapter_rank", type=int, default=None, help="bottleneck rank (overrides the reduction factor)")
    p.add_argument("--adapter_activation", default="relu", choices=["relu", "gelu"])
    p.add_argument("--kl_temp", type=float, default=3.0, help="MKD temperature (reference kl_loss default 3)")
    p.add_argument("--synthetic_batches", type=int, default=8, help="train batches per client per epoch")
    p.add_argument("--cuda_graph", action="store_true",
                   help="replay the captured train_step (feddat_b200/train/graphed.py) for full-shape batches")
    p.add_argument("--image_size", type=int, default=384)
    p.add_argument("--vit_depth", type=int, default=None, help="ALBEF: ViT blocks (default 12; smaller for smoke runs)")
    p.add_argument("--decoder_layers", type=int, default=None, help="ALBEF: answer-decoder layers (default 6)")
    p.add_argument("--text_len", type=int, default=40)
    p.add_argument("--fix_adapter0_optimizer", action="store_true",
                   help="re-enable adapter_0/adapter_1 requires_grad before each client's optimizer is built; "
                        "OFF keeps the reference behaviour where eval() leaves adapter_0 frozen (SURVEY.md F8)")
    return p


def resolve_tasks(spec):
    if spec in (None, "domain"):
        return list(DOMAIN_TASKS)                                    # main.py:358-359
    if spec.startswith("synth"):
        n = int(spec[5:]) if len(spec) > 5 and spec[5:].isdigit() else 8
        return [f"synth{i}" for i in range(n)]
    return spec.split(",")


def personal_keys(model, names):
    return [n for n in model.state_dict().keys() if any(pn in n for pn in names)]


def grad_flags(model):
    """``requires_grad`` of every parameter by name.  The reference hands each client a
    ``copy.deepcopy`` of the SERVER model (main.py:472), so every client starts from the server model's
    flags -- not from whatever the previous client's last ``train_step`` left behind
    (``set_active_adapter('adapter_0')`` switches adapter_1 off, adapter.py:71-77).  The resident model
    replaces the deepcopy, so the server's flags are recorded where the reference would copy them and
    restored before each client trains."""
    return {n: p.requires_grad for n, p in model.named_parameters()}


def restore_grad_flags(model, flags):
    for n, p in model.named_parameters():
        p.requires_grad = flags[n]


def main(argv=None, record=None):
    """``record`` (tests): a dict that receives, per round, the names of the tensors in every client's
    optimizer, the clients' adapter_1 snapshots and the averaged buffer."""
    args = build_parser().parse_args(argv)
    args.ordered_cl_tasks = resolve_tasks(args.ordered_cl_tasks)
    args.lr = args.lr if args.lr is not None else 1e-4
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(message)s", stream=sys.stderr)
    logger = logging.getLogger("feddat_b200")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    accelerator = Accelerator()
    device = accelerator.device
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl" if device.type == "cuda" else "gloo",
                                **({"device_id": device} if device.type == "cuda" else {}))
    rank = accelerator.process_index
    if not accelerator.is_main_process:
        logger.setLevel(logging.WARNING)

    torch.manual_seed(args.seed)                                     # same server model on every rank
    model = prepare_model(args, logger, device=device)
    model_config = model_configs[args.encoder_name]
    comm = FlatCommBuffer(model, model.comm_state_dict_names)
    pkeys = personal_keys(model, args.personal_params_names)
    my_clients = [(i, k) for i, k in enumerate(args.ordered_cl_tasks) if i % world == rank]
    sd = model.state_dict()
    personal = {k: {n: sd[n].detach().clone() for n in pkeys} for _, k in my_clients}      # main.py:440-450
    logger.info("clients on rank %d: %s; communicated floats: %d", rank, [k for _, k in my_clients], comm.numel)
    if not args.do_train:
        return 0

    n_clients = len(args.ordered_cl_tasks)
    server_flags = grad_flags(model)             # what deepcopy(model) would carry into every client
    for comm_round in range(args.comm_rounds):
        t0 = time.time()
        global_flat = comm.snapshot()
        client_flats = []
        for task_num, task_key in my_clients:
            with torch.no_grad():                                    # main.py:472-478 without the deepcopy
                comm.load(global_flat)
                sd = model.state_dict()
                for n in pkeys:
                    sd[n].copy_(personal[task_key][n])
            restore_grad_flags(model, server_flags)
            if args.fix_adapter0_optimizer:
                for n, p in model.named_parameters():
                    if "adapter_0" in n or "adapter_1" in n:
                        p.requires_grad = True
            out_dir = os.path.join(args.output_dir, "checkpoints", f"task{task_num}_{task_key}")
            trainer = VQATrainerSynthetic(logger, args, task_configs, model_config, device, task_key, out_dir,
                                          client_id=task_num, accelerator=accelerator)
            _, c_model = trainer.train(model, comm_round)            # main.py:485
            if record is not None:
                record.setdefault("optimizer_names", {})[(comm_round, task_key)] = list(trainer.last_optimizer_names)
            with torch.no_grad():                                    # main.py:493-503
                sd = c_model.state_dict()
                for n in pkeys:
                    personal[task_key][n].copy_(sd[n])
                client_flats.append(comm.snapshot())
            logger.info("round %d client %s: last task loss %.4f", comm_round, task_key,
                        float(trainer.last_loss) if trainer.last_loss is not None else float("nan"))
        with torch.no_grad():                                        # main.py:510, nums = [1, ...] (:455)
            get_average_net_flat(comm, client_flats, [1.0] * len(client_flats), total=float(n_clients))
        if record is not None:
            record.setdefault("client_flats", {})[comm_round] = [f.cpu() for f in client_flats]
            record.setdefault("global_flat", {})[comm_round] = comm.flat.detach().cpu().clone()
        accelerator.barrier_all_ranks()
        logger.info("round %d done in %.2f s", comm_round, time.time() - t0)

        if comm_round % 5 == 0 or args.comm_rounds - 1 == comm_round:        # main.py:520-558
            sums = torch.zeros(4, device=device)
            # the reference evaluates the SERVER model: its flags, then the eval's own side effects on them
            # (task_trainer.py:236-243 ends on set_active_adapter('adapter_1'): adapter_0 stays frozen in
            # every later round's deepcopy -- SURVEY.md F8)
            restore_grad_flags(model, server_flags)
            with torch.no_grad():
                for task_num, task_key in my_clients:
                    sd = model.state_dict()
                    for n in pkeys:
                        sd[n].copy_(personal[task_key][n])
                    trainer = VQATrainerSynthetic(logger, args, task_configs, model_config, device, task_key,
                                                  client_id=task_num, accelerator=accelerator)
                    scores = trainer.eval(model)
                    sums += torch.tensor(scores + [1.0], device=device)
            if not my_clients:           # a rank without clients still mirrors the eval's flag side effects
                model.set_active_adapter("adapter_0")
                model.set_active_adapter("adapter_1")
            server_flags = grad_flags(model)
            if world > 1:
                dist.all_reduce(sums)
            if record is not None:
                record.setdefault("eval_scores", {})[comm_round] = (sums[:3] / sums[3].clamp_min(1)).tolist()
            logger.info("Round %d: Avg test score [gating, adapter_0, adapter_1] = %s", comm_round,
                        [round(v, 2) for v in (sums[:3] / sums[3]).tolist()])
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
