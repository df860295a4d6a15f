"""ALBEF wrappers with the DAT hooks (mirror of reference src/modeling/albef.py): ``ALBEFWrapper`` (reference
:24-101), ``ALBEFContinualLearner`` (:104-183: ``set_active_adapter`` / ``activate_gating`` /
``deactivate_gating`` over the 12 ViT + 12 text-encoder + 6 text-decoder sites), ``load_albef`` (:185-249),
``create_albef_continual_learner_model`` (:252-272) and ``convert_batch_to_albef_input_dict`` (:275-286).
State-dict keys are ``albef_model.albef.{visual_encoder, text_encoder, text_decoder}...`` as upstream
(SURVEY.md Appendix B).

B200-side differences, all behaviour-preserving:
  * the wrapper also accepts PRE-TOKENISED batches (``question_ids / question_mask / answer_ids / answer_mask``
    tensors) so the tokenizer leaves the three-forward hot loop (there are no tokenizer files on
    the box); string batches go through a ``BertTokenizer`` when one is available;
  * the frozen backbone runs in bf16, adapters and the trainable LM head ('.cls.') keep fp32 masters;
  * ``defer_loss``: the train forward hands the trainer the ingredients of the answer loss instead of its value, so
    that KL + token CE + d/dlogits run as ONE fused kernel over the (answers x tokens x 30522) logits.
"""
from __future__ import annotations

import logging
from types import SimpleNamespace
from typing import Dict, List

import torch
import torch.nn as nn

from .adapter import Adapter, invalidate_packs, refresh_packs
from .albef_model import ALBEF, PAD_TOKEN_ID, SEP_TOKEN_ID
from .albef_sites import AdapterHooks


class ALBEFWrapper(nn.Module):
    """reference albef.py:24-101."""

    def __init__(self, albef: ALBEF, device, tokenizer=None):
        super().__init__()
        self.albef = albef
        self.device = device
        self.tokenizer = tokenizer
        self.defer_loss = False            # set by the trainer: fused MKD head instead of the in-model loss
        self.image_embeds = None           # set by the trainer: the visual encoder's output, computed once for two passes

    def _tok(self, texts, **kw):
        if self.tokenizer is None:
            raise RuntimeError("no BertTokenizer is available (no vocabulary files on this machine); pass pre-tokenised "
                               "question_ids / answer_ids tensors instead of strings")
        return self.tokenizer(texts, return_tensors="pt", **kw).to(self.device)

    def _images(self, batch):
        images = batch["images"].to(self.device, non_blocking=True)
        if images.dtype != torch.bfloat16 and next(self.albef.visual_encoder.parameters()).dtype == torch.bfloat16:
            images = images.to(torch.bfloat16)
        return images

    def encode_image(self, batch):
        """The visual encoder alone (albef_model.py:71), in whatever adapter mode is active."""
        return self.albef.visual_encoder(self._images(batch))

    def forward(self, batch) -> List:
        images = self._images(batch)
        pre = "question_ids" in batch
        if pre:
            question = SimpleNamespace(input_ids=batch["question_ids"].to(self.device),
                                       attention_mask=batch["question_mask"].to(self.device))
        if batch["train"]:
            weights = batch["weights"].to(self.device, non_blocking=True)
            if pre:
                answer = SimpleNamespace(input_ids=batch["answer_ids"].to(self.device),
                                         attention_mask=batch["answer_mask"].to(self.device))
            else:                                                                           # albef.py:56-57
                question = self._tok(batch["questions"], padding="longest", truncation=True, max_length=25)
                answer = self._tok(batch["answers"], padding="longest")
            idx = batch.get("answer_index")
            loss, logits = self.albef(image=images, question=question, answer=answer, train=True, alpha=batch["alpha"],
                                      k=batch["n"], weights=weights, defer_loss=self.defer_loss,
                                      answer_index=None if idx is None else idx.to(self.device, non_blocking=True),
                                      image_embeds=self.image_embeds)
            return [loss, logits]
        if pre:
            answer = SimpleNamespace(input_ids=batch["answer_list_ids"].to(self.device),
                                     attention_mask=batch["answer_list_mask"].to(self.device))
        else:                                                                               # albef.py:62-64
            question = self._tok(batch["questions"], padding="longest")
            answer = self._tok([a + "[SEP]" for a in batch["answer_list"]], padding="longest")
        topk_ids, topk_probs = self.albef(image=images, question=question, answer=answer, train=False, k=batch["k"])
        return [topk_ids, topk_probs]

    def freeze_all_weights(self):
        for p in self.albef.parameters():
            p.requires_grad = False


class ALBEFContinualLearner(nn.Module):
    """reference albef.py:104-183."""

    def __init__(self, ordered_cl_tasks: List[str], albef_model: ALBEFWrapper, task_configs: Dict):
        super().__init__()
        self.albef_model = albef_model
        self._hooks = AdapterHooks(self)

    def _adapters(self) -> List[Adapter]:
        a = self.albef_model.albef
        return ([l.output.adapter for l in a.text_encoder.encoder.layer]
                + [l.output.adapter for l in a.text_decoder.bert.encoder.layer]
                + [b.adapter for b in a.visual_encoder.blocks])

    def set_active_adapter(self, name):                               # albef.py:139-147
        for ad in self._adapters():
            ad.set_active_adapter(name)

    def deactivate_gating(self):                                      # albef.py:149-157
        for ad in self._adapters():
            ad.deactivate_gating()

    def activate_gating(self):                                        # albef.py:159-167
        for ad in self._adapters():
            ad.activate_gating()

    def get_param_adapter(self, name):
        return self._hooks.get_param_adapter(name)

    def forward(self, task_key: str, batch: Dict):                    # albef.py:169-183
        return self.albef_model(batch)

    # ------------------------------------------------------------------ B200 setup helpers / per-step hooks
    def image_forward_is_reusable(self) -> bool:
        """True when the visual encoder is a deterministic function of (image, adapter_0, adapter_2, frozen ViT): no
        active dropout in it (ALBEF's ViT is built with drop = attn_drop = drop_path = 0, albef_model.py:24-28).  The
        MKD schedule's passes A and C then share ONE ViT forward -- step B changes adapter_1 and the LM head only --
        while the BERT towers (dropout 0.1) run once per pass as in the reference."""
        vit = self.albef_model.albef.visual_encoder
        return not any(isinstance(m, nn.Dropout) and m.p > 0.0 and m.training for m in vit.modules())

    def new_step(self, train: bool = False) -> None:
        if train:
            refresh_packs(self._adapters())          # all 30 sites, both modes: ONE pack launch per train step

    def end_step(self) -> None:
        self.albef_model.image_embeds = None
        invalidate_packs(self._adapters())

    def cast_frozen_backbone(self, dtype=torch.bfloat16):
        """bf16 for everything frozen; adapters and the trainable LM head ('.cls.', main.py:127-128) stay fp32."""
        keep = ("adapter", ".cls.")
        for name, p in self.named_parameters():
            if not any(k in name for k in keep):
                p.data = p.data.to(dtype)
        return self


def load_albef(logger, model_config, checkpoint_name: str, device, pretrained_albef_name: str) -> ALBEFWrapper:
    """reference albef.py:185-249.  ``random`` builds the architecture with seeded random weights (no checkpoint
    or vocabulary files exist on the box); a real ALBEF checkpoint goes through the reference's key surgery:
    position-embedding interpolation and text_encoder layers 6-11 -> text_decoder layers 0-5."""
    import os
    tokenizer = None
    try:                                    # reference: BertTokenizer.from_pretrained('./models/bert-base-uncased')
        from transformers import BertTokenizer
        tokenizer = BertTokenizer.from_pretrained("./models/bert-base-uncased", local_files_only=True)
    except Exception:  # noqa: BLE001 - no vocabulary files on this machine: pre-tokenised batches only
        logger.warning("no BertTokenizer files under ./models/bert-base-uncased: ALBEF takes pre-tokenised batches")
    model = ALBEF(config=model_config, tokenizer=tokenizer)
    if checkpoint_name not in ("random", "random-init") and not os.path.exists(checkpoint_name):
        logger.warning("ALBEF checkpoint %r not on disk: using the random-init architecture", checkpoint_name)
    elif checkpoint_name not in ("random", "random-init"):
        from .albef_ckpt import remap_albef_checkpoint
        state = torch.load(checkpoint_name, map_location="cpu")
        state = remap_albef_checkpoint(state["model"] if "model" in state else state, model)
        missing, unexpected = model.load_state_dict(state, strict=False)
        logger.info("ALBEF checkpoint loaded: %d missing (adapters are new), %d unexpected keys", len(missing),
                    len(unexpected))
    logger.info("Successfully loaded ALBEF (%s)", checkpoint_name)
    return ALBEFWrapper(model, device, tokenizer)


def create_albef_continual_learner_model(logger, model_name_or_path: str, ordered_cl_tasks: List[str],
                                         model_config: Dict, task_configs: Dict, device):
    """reference albef.py:252-272."""
    logger = logger or logging.getLogger(__name__)
    wrapper = load_albef(logger, model_config=model_config, checkpoint_name=model_name_or_path, device=device,
                         pretrained_albef_name=model_name_or_path)
    cl_model = ALBEFContinualLearner(ordered_cl_tasks=ordered_cl_tasks, albef_model=wrapper, task_configs=task_configs)
    logger.info("Successfully created and initialized ALBEF Continual Leaner model")
    return cl_model


def convert_batch_to_albef_input_dict(batch):
    """reference albef.py:275-286 (the collate's list form), plus pre-tokenised dict batches as they are."""
    if isinstance(batch, dict):
        return dict(batch)
    return {"images": batch[0], "questions": batch[1], "answers": batch[2], "weights": batch[3], "n": batch[4],
            "alpha": batch[5]}
