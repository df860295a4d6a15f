"""``prepare_model`` for ``optimizer_mode='dat'`` (mirror of reference src/train/main.py:101-163,
248-250) plus the B200 placement step (bf16 frozen backbone, SDPA attention, fp32 adapter/head
masters on the GPU)."""
from __future__ import annotations

import logging
from types import SimpleNamespace

import torch

from ..configs.model_configs import model_configs
from ..configs.task_configs_fed import task_configs
from ..modeling.vilt import create_vilt_continual_learner_model


def prepare_model(args, logger=None, device="cuda", place=True):
    logger = logger or logging.getLogger("feddat_b200")
    albef = "albef" in args.encoder_name
    model_config = dict(model_configs[args.encoder_name])
    if "dat" not in args.optimizer_mode:
        raise NotImplementedError("only optimizer_mode='dat' is on the FedDAT hot path")

    adapter_config = {"names": [f"adapter_{i}" for i in range(3)], "device": "cpu"}       # main.py:105-112
    # the reference parses --adapter_reduction_factor but never forwards it: here it
    # is honoured, together with the new --adapter_rank / --adapter_activation
    if getattr(args, "adapter_rank", None):
        adapter_config["rank"] = args.adapter_rank
    else:
        adapter_config["adapter_reduction_factor"] = getattr(args, "adapter_reduction_factor", 16) or 16
    adapter_config["activation"] = getattr(args, "adapter_activation", "relu")
    model_config["adapter_config"] = adapter_config

    if albef:
        from ..modeling.albef import create_albef_continual_learner_model
        if getattr(args, "image_size", None):
            model_config["image_res"] = args.image_size
        for k in ("vit_depth", "decoder_layers"):                        # reduced-depth models for tests
            if getattr(args, k, None):
                model_config[k] = getattr(args, k)
        if getattr(args, "bert_overrides", None):
            model_config["bert_config"] = dict(model_config["bert_config"], **args.bert_overrides)
        model = create_albef_continual_learner_model(logger=logger, model_name_or_path=args.pretrained_model_name,
                                                     ordered_cl_tasks=args.ordered_cl_tasks, model_config=model_config,
                                                     task_configs=task_configs, device=device)
    else:
        model = create_vilt_continual_learner_model(logger=logger, model_name_or_path=args.pretrained_model_name,
                                                    ordered_cl_tasks=args.ordered_cl_tasks,
                                                    model_config=model_config, task_configs=task_configs,
                                                    device=device)
    model.comm_state_dict_names = []
    args.personal_params_names = [".cls."] if albef else ["task"]        # main.py:127-130
    for _, p in model.named_parameters():                                # main.py:138-139
        p.requires_grad = False
    if not albef:
        model.add_adapter()                                              # main.py:152-153 (ALBEF builds its sites itself)
    args.personal_params_names += ["adapter_0", "adapter_2"]             # main.py:154
    args.shared_params_names = ["adapter_1"]                             # main.py:155
    for n, p in model.named_parameters():                                # main.py:157-159
        if "adapter" in n:
            p.requires_grad = True
    for n in model.state_dict().keys():                                  # main.py:160-163
        if any(sn in n for sn in args.shared_params_names):
            model.comm_state_dict_names.append(n)
    for n, p in model.named_parameters():                                # main.py:248-250
        if "task" in n or ".cls." in n:
            p.requires_grad = True
    if place:
        place_on_gpu(model, device)
    return model


def place_on_gpu(model, device="cuda"):
    model.to(device)
    if hasattr(model, "albef_model"):
        model.albef_model.device = torch.device(device)
        model.cast_frozen_backbone(torch.bfloat16)
        return model
    model.device = torch.device(device)
    model.vilt_encoder.device = torch.device(device)
    model.cast_frozen_backbone(torch.bfloat16)
    model.vilt_encoder.enable_sdpa()
    model.vilt_encoder.enable_fused_layernorm()
    return model


def default_args(**over):
    """The flag values of reference src/train_vilt.sh:1-19 (plus the new rank/activation flags)."""
    a = dict(encoder_name="vilt", pretrained_model_name="random", optimizer_mode="dat",
             ordered_cl_tasks=["art", "abstract", "vizwiz", "toronto", "gqa"], adapter_reduction_factor=16,
             adapter_rank=None, adapter_activation="relu", adapter_config="pfeiffer", lr=1e-4, batch_size=2,
             val_batch_size=64, comm_rounds=30, local_epochs=1, num_epochs=15, seed=42, debug=0,
             do_wandb_logging=False, kl_temp=3.0)
    a.update(over)
    return SimpleNamespace(**a)
