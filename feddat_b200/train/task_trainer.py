"""Per-client trainer (mirror of reference src/train/visionlanguage_tasks/task_trainer.py).

``TaskTrainer.train`` (task_trainer.py:24-111), ``forward_pass`` (:248-264), the ``dat`` branch of
``train_step`` (:280-330, the 3-forward / 2-backward MKD schedule), ``create_optimizer``
(:477-504) and the module-level ``kl_loss`` (:506-516) keep their names, arguments and ordering of
side effects (detach points, requires_grad toggles, optimizer / scheduler stepped twice per batch,
``zero_grad`` -> None grads).  The loss arithmetic runs in the fused sm_100a MKD kernel.
"""
from __future__ import annotations

import contextlib
import os
from typing import Dict

import torch
import torch.nn as nn
from torch.optim import AdamW

from .. import ops


class _MkdLossFunction(torch.autograd.Function):
    """(task_weight * task + kl_weight * kl) with d/dlogits from the same kernel launch."""

    @staticmethod
    def forward(ctx, logits, teacher, target, temp, kl_weight, task_weight):
        shape = logits.shape
        lg = logits.contiguous().float()
        loss3, dlogits = ops.mkd_loss(lg, teacher.contiguous().float(),
                                      None if target is None else target.contiguous().float(),
                                      temp, kl_weight, task_weight, need_grad=True)
        ctx.save_for_backward(dlogits)
        ctx.shape, ctx.in_dtype = shape, logits.dtype
        ctx.mark_non_differentiable(loss3)
        return loss3[0], loss3

    @staticmethod
    def backward(ctx, g_total, _g3):
        (dlogits,) = ctx.saved_tensors
        g = (dlogits * g_total).view(ctx.shape)
        return g.to(ctx.in_dtype), None, None, None, None, None


def _nvtx(name: str):
    """NVTX range around a phase of the MKD schedule (visible in nsys / ncu --nvtx timelines).
    A no-op without CUDA."""
    if torch.cuda.is_available():
        return torch.cuda.nvtx.range(f"feddat/{name}")
    return contextlib.nullcontext()


class _MkdCeFunction(torch.autograd.Function):
    """ALBEF objective (task_trainer.py:296-301 with the answer loss of albef_model.py:142-143) as ONE fused
    launch over the decoder's prediction scores: value and d/dscores (feddat_mkd_ce_loss)."""

    @staticmethod
    def forward(ctx, scores, teacher, labels, seq_weight, temp, kl_weight, task_weight):
        loss3, dscores = ops.mkd_ce_loss(scores.contiguous(), teacher, labels, seq_weight, temp, kl_weight, task_weight,
                                         need_grad=True)
        ctx.save_for_backward(dscores)
        ctx.mark_non_differentiable(loss3)
        return loss3[0], loss3

    @staticmethod
    def backward(ctx, g_total, _g3):
        (dscores,) = ctx.saved_tensors
        return dscores * g_total.to(dscores.dtype), None, None, None, None, None, None


def kl_loss(output, target, temp=3):
    """task_trainer.py:506-516: T^2 * KL(softmax(target/T) || softmax(output/T)), 'batchmean'."""
    total, _ = _MkdLossFunction.apply(output, target, None, float(temp), 1.0, 0.0)
    return total


def mkd_objective(logits, teacher, target, temp=3):
    """(BCEWithLogits('mean')(logits, target) * C + kl_loss(logits, teacher.detach(), temp)) / 2
    (task_trainer.py:299-301) in one fused launch.  Returns (L, loss3 = [L, kl, task])."""
    return _MkdLossFunction.apply(logits, teacher.detach(), target, float(temp), 0.5, 0.5)


def get_polynomial_decay_schedule_with_warmup(optimizer, num_warmup_steps, num_training_steps, lr_end=0.0,
                                              power=1.0):
    from transformers import get_polynomial_decay_schedule_with_warmup as hf
    return hf(optimizer, num_warmup_steps=num_warmup_steps, num_training_steps=num_training_steps,
              lr_end=lr_end, power=power)


DEFER_WGRAD = True        # batched schedule: all sites' weight gradients in ONE launch after the backward (ops.DeferredWgrad)


def _deferred():
    return ops.deferred_wgrad() if DEFER_WGRAD else contextlib.nullcontext()


# The task head (fp32 masters: Linear 768 -> 1536, LayerNorm, GELU, Linear 1536 -> labels on 32 / 64 pooled rows) runs its
# matmuls on TF32 tensor cores inside a train step: fp32 storage and accumulation, 10-bit-mantissa products -- the
# precision of the reference's own head under accelerate's fp16 autocast (task_trainer.py:50-63), above the bf16 of the
# rest of the step.  As fp32 SIMT GEMMs the ~13 launch-bound head matmuls cost 0.09 ms per step (scripts/ab_step.py:
# 6.21 -> 6.11 ms).  False = strict fp32 matmuls.
HEAD_TF32 = True
USE_OWN_ADAMW = True      # csrc/adamw.cu behind train/fused_adamw.py; False = torch.optim.AdamW(fused, capturable)


def split_step(optimizer, scheduler, first_params):
    """All gradients of the batched schedule are present at once; apply them as the reference's two
    optimizer steps would: ``first_params`` (adapter_1) with the CURRENT learning rate, scheduler step,
    every other parameter with the next learning rate, scheduler step, zero_grad.  AdamW skips parameters
    whose ``.grad`` is None, and its state (moments, step count) is per parameter."""
    first = {id(p) for p in first_params}
    held = []
    for group in optimizer.param_groups:
        for p in group["params"]:
            if p.grad is not None and id(p) not in first:
                held.append((p, p.grad))
                p.grad = None
    optimizer.step()
    if scheduler is not None:
        scheduler.step()
    for p in first_params:
        p.grad = None
    for p, g in held:
        p.grad = g
    optimizer.step()
    if scheduler is not None:
        scheduler.step()
    optimizer.zero_grad()


class TaskTrainer(nn.Module):

    def __init__(self, **kwargs):
        super().__init__()
        self.kl_criterion = kl_loss                      # task_trainer.py:15
        self.kl_temp = 3                                 # kl_loss default; BASELINE cfg1 uses 2.0
        # passes A and C of the MKD schedule run the SAME encoder function (gating: adapter_0 +
        # adapter_2 + frozen backbone; step B only updates adapter_1 and the head) whenever the
        # encoder has no dropout (ViLT): one grad-enabled forward serves both
        self.reuse_gating_forward = True
        # Batched schedule (same arithmetic, see _train_step_batched): the gating pass and the adapter_1 pass
        # run as ONE row-stacked forward and ONE row-stacked backward through the frozen backbone
        self.batched_passes = True
        # tests: ``grad_probe(tag, model)`` is called with every gradient of a pass in place, BEFORE the
        # optimizer consumes it (tags "B" / "C" in the reference order; "B_head" / "BC" in the batched one)
        self.grad_probe = None

    def _probe(self, tag, model):
        if self.grad_probe is not None:
            self.grad_probe(tag, model)

    # ------------------------------------------------------------------ train (task_trainer.py:24-111)
    def train(self, model, er=None, ewc=None, der=None, derpp=None, pnn=None, hat=None):
        sd = model.state_dict()
        for name in sd.keys():                           # :36-41  adapter_2 <- adapter_1
            if "adapter_1" in name:
                name_tgt = name.replace("adapter_1", "adapter_2")
                if name_tgt in sd:
                    sd[name_tgt].data.copy_(sd[name].data)
        for n, p in model.named_parameters():            # :43-45
            if "adapter_2" in n:
                p.requires_grad = False
        if getattr(self, "task_output_dir", None) and not os.path.isdir(self.task_output_dir):
            os.makedirs(self.task_output_dir, exist_ok=True)

        model = self.accelerator.prepare(model)
        optimizer = self.create_optimizer(model, self.args.optimizer_mode)
        ids = {id(p) for g in optimizer.param_groups for p in g["params"]}
        self.last_optimizer_names = [n for n, p in model.named_parameters() if id(p) in ids]
        scheduler = get_polynomial_decay_schedule_with_warmup(
            optimizer, num_warmup_steps=int(self.max_steps * self.warmup_ratio),
            num_training_steps=self.max_steps, lr_end=0, power=1)
        model.zero_grad()
        optimizer, scheduler = self.accelerator.prepare(optimizer, scheduler)
        loader = self.vqa_train_dataloader

        loss = None
        graphed = None          # --cuda_graph: replay the captured step for every full-shape batch
        for epoch in range(self.local_epochs):
            model.train()
            for step, batch in enumerate(loader):
                if getattr(self.args, "debug", 0) > 0 and step > self.args.debug:
                    break
                if "vilt" not in self.args.encoder_name:
                    batch = self.add_alpha(epoch, batch, step)
                if getattr(self.args, "cuda_graph", False) and isinstance(batch, dict):
                    from .graphed import GraphedDictStep, GraphedTrainStep
                    if graphed is None and "encodings" in batch:
                        graphed = GraphedTrainStep(self, model, optimizer, scheduler, batch, warmup=1)
                    elif graphed is None and "answer_index" in batch and all(
                            v.is_cuda for v in batch.values() if isinstance(v, torch.Tensor)):
                        # ALBEF: pre-tokenised device batches with the answer -> question map as a tensor
                        graphed = GraphedDictStep(self, model, optimizer, scheduler, batch, warmup=1)
                    if graphed is not None and graphed.accepts(batch):
                        loss = graphed(batch)
                        continue
                loss = self.train_step(model, step, batch, optimizer, scheduler, hooks=None, epoch=epoch)
        self.accelerator.wait_for_everyone()
        self.last_loss = loss
        del optimizer, scheduler
        unwrapped_model = self.accelerator.unwrap_model(model)
        self.accelerator.free_memory()
        return 0.0, unwrapped_model

    # ------------------------------------------------------------------ forward_pass (:248-264)
    def forward_pass(self, model, batch, do_eval: bool = False):
        inputs = self.batch2inputs_converter(batch)
        if "albef" in self.args.encoder_name and not do_eval:
            inputs["train"] = True
        vilt_like = self.args.encoder_name in ["vilt", "viltbert"]
        if do_eval is True:
            with torch.no_grad():
                return model(task_key=self.task_key, **inputs) if vilt_like else model(self.task_key, inputs)
        return model(task_key=self.task_key, **inputs) if vilt_like else model(self.task_key, inputs)

    # ------------------------------------------------------------------ train_step (:266-330)
    def _objective(self, logits, teacher, target, task_loss):
        """(L, task) with L = (task + kl(logits, teacher.detach())) / 2.  Fused BCE+KL launch for the ViLT criterion
        (BCEWithLogits 'mean' * C, train_vqa_crossvqa.py:237); KL kernel + given task loss otherwise."""
        if hasattr(task_loss, "prediction_scores"):
            # ALBEF with the deferred answer loss (modeling/albef_model.py LazyAnswerLoss): KL + weighted token CE
            # + gradient in one pass over the [answers, tokens, 30522] scores
            lazy = task_loss
            total, loss3 = _MkdCeFunction.apply(lazy.prediction_scores, teacher.detach(), lazy.labels,
                                                lazy.weights.float() / lazy.batch_size, float(self.kl_temp), 0.5, 0.5)
            return total, loss3[2]
        fused = (task_loss is None and isinstance(self.loss_criterion, nn.BCEWithLogitsLoss)
                 and self.loss_criterion.reduction == "mean" and self.kl_criterion is kl_loss
                 and self.loss_criterion.weight is None and self.loss_criterion.pos_weight is None)
        if fused:
            total, loss3 = mkd_objective(logits, teacher, target, self.kl_temp)
            return total, loss3[2]
        if task_loss is None:
            task_loss = self.loss_criterion(logits, target) * target.shape[1]
        if self.kl_criterion is kl_loss:
            loss_kl = kl_loss(logits, teacher.clone().detach(), self.kl_temp)
        else:
            loss_kl = self.kl_criterion(logits, teacher.clone().detach())
        return (task_loss + loss_kl) / 2, task_loss.detach()

    def _objective_is_fused(self) -> bool:
        return (isinstance(self.loss_criterion, nn.BCEWithLogitsLoss) and self.loss_criterion.reduction == "mean"
                and self.kl_criterion is kl_loss and self.loss_criterion.weight is None
                and self.loss_criterion.pos_weight is None)

    def _train_step_batched(self, model, batch, target, optimizer, scheduler):
        """The dat branch of train_step (task_trainer.py:280-330) with passes A/C and B batched.

        The reference runs  A: logits_all (gating, no grad) | B: forward adapter_1, L_1, backward, step |
        C: forward gating, L_0, backward, step.  Facts used: (i) ViLT has no dropout, so C's encoder output
        equals A's (step B changes neither adapter_0, adapter_2 nor the backbone); (ii) AdamW updates every
        parameter independently, so "step B" may be split into "task head now, adapter_1 later" as long as
        each parameter is stepped with the same gradient, the same learning rate and the same step count;
        (iii) backward through the frozen backbone is linear in the incoming gradient rows.  Hence:

          1. ONE forward over the batch stacked twice: rows [0, B) through the gating pair, rows [B, 2B)
             through adapter_1 (Adapter.set_dual) -> enc_A, enc_B
          2. logits_all = head(enc_A) (old head, no grad); L_1 from head(enc_B); backward through the HEAD
             only (-> head grads, d enc_B); optimizer step of the head with the first learning rate
          3. L_0 from the UPDATED head(enc_A) and logits_1; backward through the head (-> head grads, d enc_A)
          4. ONE backward through the encoder with [d enc_A; d enc_B] -> adapter_0 and adapter_1 grads
          5. step adapter_1 (first learning rate), scheduler step, step adapter_0 + head (second learning
             rate), scheduler step

        Every parameter receives the updates of the reference schedule; the frozen backbone runs once
        forward and once backward over 2B rows instead of twice / twice over B rows (fuller GEMM waves,
        half the launches).  Checked against the reference trainer by tests/test_train_step_gpu.py."""
        inner = model.module
        with _nvtx("batched_fwd_A+B"):
            enc = inner.encode_dual(**self.batch2inputs_converter(batch))
        b = enc.shape[0] // 2
        leaf = enc.detach().requires_grad_(True)                     # head gradients stop here until step 4
        enc_a, enc_b = leaf[:b], leaf[b:]
        # (A) and (B) read the SAME (old) head: one call over both halves -- the head is ~25 tiny fp32 launches per
        # call, launch-bound at batch 32.  Only logits_1 carries a gradient: the A rows are detached going in (their
        # rows of every head-internal gradient are exact zeros, so the head's parameter gradients are those of pass B)
        both = inner.classify(self.task_key, torch.cat([enc_a.detach(), enc_b], dim=0))
        logits_all, logits_1 = both[:b].detach(), both[b:]           # (A): old head, gating encoder;  (B)
        L_1, _ = self._objective(logits_1, logits_all, target, None)
        self.accelerator.backward(L_1)                               # head grads + d enc_B (adapters: none yet)
        self._probe("B_head", model)
        optimizer.step()                                             # only the head has gradients
        optimizer.zero_grad()

        logits_0 = inner.classify(self.task_key, enc_a)              # (C): updated head
        L_0, loss_0 = self._objective(logits_0, logits_1, target, None)
        self.accelerator.backward(L_0)                               # head grads + d enc_A
        # the sites' weight gradients are not needed before step 5: the backward queues them and ONE launch at
        # its end computes all of them (ops.DeferredWgrad)
        with _nvtx("batched_bwd_C+B"), _deferred():
            enc.backward(leaf.grad)                                  # adapter_0 (rows A) and adapter_1 (rows B)
        self._probe("BC", model)

        a1 = getattr(self, "_a1_params", None)
        if a1 is None or a1[0] is not model:
            a1 = self._a1_params = (model, [p for n, p in model.named_parameters() if "adapter_1" in n])
        split_step(optimizer, scheduler, a1[1])

        inner.activate_gating()                                      # leave the modes as the reference does
        inner.set_active_adapter("adapter_0")
        self.last_logits = (logits_all.detach(), logits_1.detach(), logits_0.detach())
        self.last_objectives = (L_1.detach(), L_0.detach())
        return loss_0

    def train_step(self, model, step, batch, optimizer=None, scheduler=None, hooks=None, epoch=None):
        target = None
        if isinstance(batch, dict) and "target_scores" in batch.keys():
            target = batch["target_scores"].to(self.device)
        if "dat" not in self.args.optimizer_mode:
            raise NotImplementedError("only optimizer_mode='dat' is on the FedDAT hot path")
        albef = "albef" in self.args.encoder_name

        inner = model.module
        if hasattr(inner, "new_step"):
            inner.new_step(train=True)          # per-step caches; all sites' bf16 operands packed in one launch
        tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32 or HEAD_TF32
        try:
            return self._train_step_dat(model, inner, batch, target, optimizer, scheduler, albef)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = tf32
            if hasattr(inner, "end_step"):
                inner.end_step()

    def _train_step_dat(self, model, inner, batch, target, optimizer, scheduler, albef):
        if albef and hasattr(inner, "albef_model"):
            # fused MKD head whenever the KL criterion is the reference's own kl_loss
            inner.albef_model.defer_loss = self.kl_criterion is kl_loss
        reuse = (self.reuse_gating_forward and not albef
                 and getattr(inner, "gating_forward_is_reusable", lambda: False)())
        img = None                # ALBEF: the gating ViT forward shared by passes A and C
        if (reuse and self.batched_passes and optimizer is not None and hasattr(inner, "encode_dual")
                and isinstance(batch, dict) and batch.get("encodings", {}).get("dense_masks", False)
                and self._objective_is_fused() and 2 * inner._adapters()[0].rank <= ops.MAX_R_TOTAL):
            return self._train_step_batched(model, batch, target, optimizer, scheduler)
        if reuse:                                                    # (A) + encoder half of (C)
            inner.activate_gating()
            inner.set_active_adapter("adapter_0")
            enc_all = inner.encode(**self.batch2inputs_converter(batch))
            with torch.no_grad():
                logits_all = inner.classify(self.task_key, enc_all)
        else:
            # ALBEF: passes A and C see the same image through the same gating adapters and its ViT has no dropout, so
            # ONE grad-enabled ViT forward serves both (3 -> 2 ViT forwards per step; the BERT towers have dropout
            # and run once per pass).  The modes are set as pass C sets them (:311-312).
            if (albef and self.reuse_gating_forward and hasattr(inner, "albef_model")
                    and getattr(inner, "image_forward_is_reusable", lambda: False)()):
                inner.activate_gating()
                inner.set_active_adapter("adapter_0")
                with _nvtx("pass_AC_vit_fwd"):
                    img = inner.albef_model.encode_image(self.batch2inputs_converter(batch))
                inner.albef_model.image_embeds = img.detach()
            with torch.no_grad(), _nvtx("pass_A_gating_nograd"):     # (A) :283-287
                model.module.activate_gating()
                _, logits_all = self.forward_pass(model, batch, do_eval=False)
            if img is not None:
                inner.albef_model.image_embeds = None

        model.module.deactivate_gating()                             # (B) :290-308
        model.module.set_active_adapter("adapter_1")
        with _nvtx("pass_B_adapter1_fwd"):
            output_1_0, logits_1 = self.forward_pass(model, batch, do_eval=False)
            L_1, _ = self._objective(logits_1, logits_all, target, output_1_0 if albef else None)
        with _nvtx("pass_B_adapter1_bwd"), _deferred():      # all sites' weight gradients after the backward pass
            self.accelerator.backward(L_1)
        self._probe("B", model)
        if optimizer is not None:
            optimizer.step()
            if scheduler is not None:
                scheduler.step()
            optimizer.zero_grad()

        model.module.activate_gating()                               # (C) :311-328
        model.module.set_active_adapter("adapter_0")
        if reuse:       # same pooled output as a fresh forward (bit-identical), head re-applied after step B
            output_0_0, logits_0 = enc_all, inner.classify(self.task_key, enc_all)
        else:
            if img is not None:
                inner.albef_model.image_embeds = img                 # pass A's ViT output, with its graph
            output_0_0, logits_0 = self.forward_pass(model, batch, do_eval=False)
            if img is not None:
                inner.albef_model.image_embeds = None
        L_0, loss_0 = self._objective(logits_0, logits_1, target, output_0_0 if albef else None)
        with _nvtx("pass_C_gating_bwd"), _deferred():
            self.accelerator.backward(L_0)
        self._probe("C", model)
        if optimizer is not None:
            optimizer.step()
            if scheduler is not None:
                scheduler.step()
            optimizer.zero_grad()
        self.last_logits = (logits_all.detach(), logits_1.detach(), logits_0.detach())
        self.last_objectives = (L_1.detach(), L_0.detach())
        return loss_0          # the task term, as the reference does (task_trainer.py:319,330)

    # ------------------------------------------------------------------ create_optimizer (:477-504)
    def create_optimizer(self, model, mode="full"):
        no_decay = ["bias", "LayerNorm.weight"]
        groups = [
            {"params": [p for n, p in model.named_parameters()
                        if (not any(nd in n for nd in no_decay)) and p.requires_grad],
             "weight_decay": self.weight_decay},
            {"params": [p for n, p in model.named_parameters()
                        if (any(nd in n for nd in no_decay)) and p.requires_grad],
             "weight_decay": 0.0},
        ]
        dev = next((p.device for g in groups for p in g["params"] if p.is_cuda), None)
        if dev is None:
            return AdamW(groups, lr=self.lr, eps=self.adam_epsilon, betas=(0.9, 0.98))
        # fused + capturable with a device-tensor lr: same arithmetic, and the two optimizer steps of
        # a train step can live inside a CUDA graph (feddat_b200/train/graphed.py)
        if USE_OWN_ADAMW and all(p.dtype == torch.float32 and p.is_cuda for g in groups for p in g["params"]):
            from .fused_adamw import FusedAdamW           # same rule and state, this repo's multi-tensor kernel
            opt = FusedAdamW(groups, lr=float(self.lr), eps=self.adam_epsilon, betas=(0.9, 0.98))
        else:
            opt = AdamW(groups, lr=float(self.lr), eps=self.adam_epsilon, betas=(0.9, 0.98), fused=True, capturable=True)
        # per-group lr tensors are installed AFTER construction: ``defaults["lr"]`` must stay a float,
        # the HF polynomial schedule divides by it (lr_init) -- if it aliased the live lr tensor, the
        # warm-up step that sets lr = 0 would turn every later factor into 0/0
        for g in opt.param_groups:
            g["lr"] = torch.tensor(float(g["lr"]), device=dev)
        return opt
