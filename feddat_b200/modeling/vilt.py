"""ViLT wrappers with the DAT injection hooks (mirror of reference src/modeling/vilt.py).

Same class / method surface as the reference: ``ViltEncoderWrapper`` (processor + HF ``ViltModel``),
``ViltContinualLearner`` (per-task MLP heads vilt.py:200-210, ``add_adapter`` /
``set_active_adapter`` / ``activate_gating`` / ``deactivate_gating`` / ``get_param_adapter``
vilt.py:356-382), ``create_vilt_continual_learner_model`` (vilt.py:421-452) and
``convert_batch_to_vilt_input_dict`` (vilt.py:455-459).  State-dict keys are identical
(``vilt_encoder.vilt.encoder.layer.{i}.output.adapter.adapter_{j}_{down,up}.{weight,bias}``,
``task_layer.{task}.clf_*``), so the federated round loop's substring selection works unchanged.

B200-side differences, all behaviour-preserving:
  * forward also accepts PRE-ENCODED tensors (``input_ids, attention_mask, token_type_ids,
    pixel_values, pixel_mask``) so the tokenizer / image processor leave the three-forward hot loop
    (SURVEY.md F10); PIL/str inputs still go through ``process_inputs`` when a processor exists.
  * a dense fast path for batches whose masks are all ones: patch order is kept instead of the HF
    ``torch.multinomial`` shuffle (the encoder is permutation-equivariant and the pooled CLS output
    is invariant to patch order), no host syncs, SDPA attention, and the frozen embedding output is
    computed once per batch and reused by the three MKD passes (ViLT dropout is 0.0).
  * the frozen backbone runs in bf16; adapters and task heads keep fp32 master weights.
"""
from __future__ import annotations

import logging
import math
import types
from collections import OrderedDict
from typing import Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .adaptered_output import Adaptered_ViltOutput

ENCODING_KEYS = ("input_ids", "attention_mask", "token_type_ids", "pixel_values", "pixel_mask")


class _FrozenQKV(torch.autograd.Function):
    """q, k, v = three frozen Linear projections of one tensor.  Forward is the stock three GEMMs; backward
    accumulates the three data gradients INSIDE the GEMMs (``addmm_`` with beta = 1) instead of autograd's
    three GEMMs + two gradient-sum kernels (44 elementwise launches per train step)."""

    @staticmethod
    def forward(ctx, x, wq, bq, wk, bk, wv, bv):
        ctx.save_for_backward(wq, wk, wv)
        return F.linear(x, wq, bq), F.linear(x, wk, bk), F.linear(x, wv, bv)

    @staticmethod
    def backward(ctx, gq, gk, gv):
        wq, wk, wv = ctx.saved_tensors
        shape = gq.shape
        d_out = shape[-1]
        dx = torch.mm(gq.reshape(-1, d_out), wq)
        dx.addmm_(gk.reshape(-1, d_out), wk)
        dx.addmm_(gv.reshape(-1, d_out), wv)
        return dx.view(*shape[:-1], wq.shape[1]), None, None, None, None, None, None


FUSE_ATTENTION = True      # own short-sequence attention kernels (csrc/attn.cu, attn_bwd.cu) where they apply


class _FrozenQKVAttention(torch.autograd.Function):
    """ctx = softmax(q k^T / sqrt(d)) v with q, k, v the three FROZEN projections of x, for short sequences:
    ONE [M, 768] x [768, 2304] GEMM (concatenated frozen weights), this repo's attention kernel on column slices of
    its output (feddat_attn_fwd), and in backward feddat_attn_bwd writing dq | dk | dv into ONE [M, 2304] tensor that
    a single GEMM against the concatenated weight turns into dx -- instead of 3 GEMMs + cuDNN SDPA forward and
    3 cuDNN launches + 3 GEMMs backward (HF ViltSelfAttention; reference src/modeling/vilt.py:19,127)."""

    @staticmethod
    def forward(ctx, x, w_qkv, b_qkv, heads, scale):
        from .. import ops
        b, s, d = x.shape
        qkv = torch.addmm(b_qkv, x.reshape(-1, d), w_qkv.t()).view(b, s, 3, heads, d // heads)
        o, lse = ops.attn_fwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], scale)
        ctx.save_for_backward(qkv, o, lse, w_qkv)
        ctx.scale = scale
        return o.view(b, s, d)

    @staticmethod
    def backward(ctx, g):
        from .. import ops
        qkv, o, lse, w_qkv = ctx.saved_tensors
        b, s, _, heads, dh = qkv.shape
        g = g.contiguous().view(b, s, heads, dh)
        dqkv = ops.attn_bwd(g, qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], o, lse, ctx.scale, packed=True)
        dx = torch.mm(dqkv.view(b * s, 3 * heads * dh), w_qkv)
        return dx.view(b, s, heads * dh), None, None, None, None


def _qkv_concat(att):
    """[2304, 768] weight and [2304] bias of the three frozen projections, cached on the attention module and
    rebuilt when any of the six tensors changes (load_state_dict, dtype casts)."""
    ps = (att.query.weight, att.key.weight, att.value.weight, att.query.bias, att.key.bias, att.value.bias)
    key = tuple((p.data_ptr(), p._version, p.dtype) for p in ps)
    cached = getattr(att, "_feddat_qkv", None)
    if cached is None or cached[0] != key:
        with torch.no_grad():
            cached = (key, torch.cat([p.detach() for p in ps[:3]], 0).contiguous(),
                      torch.cat([p.detach() for p in ps[3:]], 0).contiguous())
        att._feddat_qkv = cached
    return cached[1], cached[2]


def _own_attention_ok(att, x, frozen, needs_grad) -> bool:
    from .. import ops
    dh = att.attention_head_size
    return (FUSE_ATTENTION and frozen and x.is_cuda and x.dtype == torch.bfloat16 and att.query.weight.dtype == torch.bfloat16
            and dh == 64 and x.dim() == 3 and x.shape[1] <= (ops.ATTN_MAX_S_BWD if needs_grad else ops.ATTN_MAX_S_FWD)
            and (att.dropout.p == 0.0 or not att.training))


def _sdpa_self_attention_forward(self, hidden_states, attention_mask=None, output_attentions=False):
    """softmax(QK^T / sqrt(d)) V of HF ViltSelfAttention: this repo's short-sequence kernels when they apply
    (``_FrozenQKVAttention``), torch SDPA otherwise (frozen backbone op)."""
    if output_attentions:
        from .fused_ln import _fallback_once
        _fallback_once("_sdpa_self_attention_forward", "attention maps requested: explicit softmax path")
        return type(self).forward(self, hidden_states, attention_mask, output_attentions)
    b, s, _ = hidden_states.shape
    h, dh = self.num_attention_heads, self.attention_head_size
    mods = (self.query, self.key, self.value)
    frozen = not any(p.requires_grad for m in mods for p in m.parameters()) and all(m.bias is not None for m in mods)
    if attention_mask is None and _own_attention_ok(self, hidden_states, frozen,
                                                        hidden_states.requires_grad and torch.is_grad_enabled()):
        w_qkv, b_qkv = _qkv_concat(self)
        return (_FrozenQKVAttention.apply(hidden_states.contiguous(), w_qkv, b_qkv, h, 1.0 / math.sqrt(dh)),)
    if FUSE_ATTENTION and attention_mask is None:
        from .fused_ln import _fallback_once
        _fallback_once("_sdpa_self_attention_forward", "sequence longer than the short-sequence attention kernels cover, "
                       "non-bf16 / unfrozen projections or active dropout: torch SDPA")
    if frozen and hidden_states.requires_grad:
        q, k, v = _FrozenQKV.apply(hidden_states, self.query.weight, self.query.bias, self.key.weight,
                                   self.key.bias, self.value.weight, self.value.bias)
    else:
        q, k, v = self.query(hidden_states), self.key(hidden_states), self.value(hidden_states)
    q = q.view(b, s, h, dh).transpose(1, 2)
    k = k.view(b, s, h, dh).transpose(1, 2)
    v = v.view(b, s, h, dh).transpose(1, 2)
    if attention_mask is not None:
        attention_mask = attention_mask.to(q.dtype)
    ctx = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask,
                                         dropout_p=self.dropout.p if self.training else 0.0)
    return (ctx.transpose(1, 2).reshape(b, s, h * dh),)


class ViltEncoderWrapper(nn.Module):
    """reference vilt.py:22-148."""

    def __init__(self, processor, vilt, device):
        super().__init__()
        self.vilt = vilt
        self.device = device
        self.processor = processor
        self.max_text_length = self.vilt.config.max_position_embeddings
        self.encoder_dim = self.vilt.config.hidden_size
        self.dense_fast_path = True
        self._embed_cache = None
        self.expand_modality_type_embeddings()

    def expand_modality_type_embeddings(self, type_vocab_size=3):
        """vilt.py:102-113 (the reference constructor always runs this, vilt.py:45)."""
        self.vilt.config.modality_type_vocab_size = type_vocab_size
        emb_data = self.vilt.embeddings.token_type_embeddings.weight.data
        new = nn.Embedding(type_vocab_size, self.encoder_dim).to(emb_data.device, emb_data.dtype)
        new.weight.data[0, :] = emb_data[0, :]
        new.weight.data[1, :] = emb_data[1, :]
        new.weight.data[2, :] = emb_data[1, :]
        self.vilt.embeddings.token_type_embeddings = new

    def process_inputs(self, images: List, texts: List[str]) -> Dict:
        """vilt.py:87-100."""
        if self.processor is None:
            raise RuntimeError("no ViltProcessor is available (no pretrained files on this machine); "
                               "pass pre-encoded tensors instead of images/texts")
        return self.processor(images=images, text=texts, max_length=self.max_text_length, padding=True,
                              truncation=True, return_tensors="pt").to(self.device)

    def enable_sdpa(self):
        for layer in self.vilt.encoder.layer:
            att = layer.attention.attention
            att.forward = types.MethodType(_sdpa_self_attention_forward, att)

    def enable_fused_layernorm(self):
        """Fused (residual add +) LayerNorm kernels for the frozen blocks (modeling/fused_ln.py)."""
        from . import fused_ln
        fused_ln.enable(self.vilt)
        self.fused_layernorm = True

    # ------------------------------------------------------------------ dense fast path
    def _dense_embeddings(self, input_ids, token_type_ids, pixel_values):
        emb = self.vilt.embeddings
        cfg = self.vilt.config
        text = emb.text_embeddings(input_ids=input_ids, token_type_ids=token_type_ids)
        # patch embedding = Conv2d(3, 768, kernel 32, stride 32): non-overlapping patches, so it is ONE
        # [B*h*w, 3072] x [3072, 768] GEMM (cuDNN's generic fprop + NCHW<->NHWC transposes took
        # 0.86 ms per step for this 22 GFLOP product, profiles/r1_summary.md)
        proj = emb.patch_embeddings.projection
        ps = proj.kernel_size[0]
        b, cin, hh, ww = pixel_values.shape
        h, w = hh // ps, ww // ps
        if (pixel_values.is_cuda and proj.weight.dtype == torch.bfloat16 and pixel_values.is_contiguous()
                and pixel_values.dtype in (torch.float32, torch.bfloat16) and ps % 8 == 0 and ww % 8 == 0):
            from .. import ops
            patches = ops.patchify(pixel_values, ps)                  # cut + cast in one pass (csrc/patchify.cu)
        else:
            px = pixel_values.to(proj.weight.dtype)[:, :, :h * ps, :w * ps]
            patches = px.reshape(b, cin, h, ps, w, ps).permute(0, 2, 4, 1, 3, 5).reshape(b * h * w, cin * ps * ps)
        x = F.linear(patches, proj.weight.view(proj.out_channels, -1), proj.bias).view(b, h * w, -1)
        c = x.shape[-1]
        pd = cfg.image_size // cfg.patch_size
        spatial = emb.position_embeddings[:, 1:, :].transpose(1, 2).view(1, c, pd, pd)
        if (h, w) != (pd, pd):
            spatial = F.interpolate(spatial.float(), size=(h, w), mode="bilinear", align_corners=True).to(x.dtype)
        x = x + spatial.flatten(2).transpose(1, 2)
        cls = emb.cls_token.expand(b, -1, -1) + emb.position_embeddings[:, 0, :][:, None, :]
        img = emb.dropout(torch.cat((cls, x), dim=1))
        tt = emb.token_type_embeddings.weight
        return torch.cat([text + tt[0], img + tt[1]], dim=1)

    def _dense_forward(self, input_ids, token_type_ids, pixel_values):
        # The cache only serves the passes of ONE train step (TaskTrainer clears it at the start of every
        # train / eval step): ``_version`` does not move for ``.data.copy_`` or external writes into a static
        # input buffer, so the key alone could not tell a refilled buffer from the old batch.
        key = (input_ids.data_ptr(), input_ids._version, pixel_values.data_ptr(), pixel_values._version,
               tuple(pixel_values.shape),
               None if token_type_ids is None else (token_type_ids.data_ptr(), token_type_ids._version))
        if self._embed_cache is not None and self._embed_cache[0] == key:
            hidden = self._embed_cache[1]
        else:
            with torch.no_grad():                       # embeddings are frozen (main.py:138-139)
                hidden = self._dense_embeddings(input_ids, token_type_ids, pixel_values)
            # keep references to the inputs so their storage (and the key) stays valid
            self._embed_cache = (key, hidden, input_ids, pixel_values, token_type_ids)
        for layer in self.vilt.encoder.layer:
            hidden = layer(hidden, None)[0]
        # the pooler reads only the CLS row of the final LayerNorm (per-row op): normalise that row alone
        return self.vilt.pooler(self.vilt.layernorm(hidden[:, 0:1]))

    def _dense_forward_stacked(self, input_ids, token_type_ids, pixel_values):
        """Two copies of the batch stacked along the batch dimension through ONE pass of the frozen layers
        (the adapters are in dual mode: each half sees its own DAT branch).  Returns (2 * batch, hidden)."""
        with torch.no_grad():
            hidden = self._dense_embeddings(input_ids, token_type_ids, pixel_values)
            hidden = torch.cat([hidden, hidden], dim=0)
        for layer in self.vilt.encoder.layer:
            hidden = layer(hidden, None)[0]
        return self.vilt.pooler(self.vilt.layernorm(hidden[:, 0:1]))

    def forward(self, **encodings: Dict) -> torch.FloatTensor:
        """vilt.py:115-129: returns ``pooler_output`` (batch, hidden)."""
        dense = encodings.pop("dense_masks", False)
        if dense and self.dense_fast_path:
            return self._dense_forward(encodings["input_ids"], encodings.get("token_type_ids"),
                                       encodings["pixel_values"])
        output = self.vilt(**encodings)
        return output.pooler_output

    def freeze_all_weights(self):
        for p in self.vilt.parameters():
            p.requires_grad = False

    def freeze_bottom_k_layers(self, k: int):
        assert k < len(self.vilt.encoder.layer)
        for p in self.vilt.embeddings.parameters():
            p.requires_grad = False
        for i in range(k):
            for p in self.vilt.encoder.layer[i].parameters():
                p.requires_grad = False


class ViltContinualLearner(nn.Module):
    """reference vilt.py:152-382 (classification, single-image tasks: the VQA path of train_vilt.sh)."""

    def __init__(self, ordered_cl_tasks: List[str], encoder: ViltEncoderWrapper, encoder_dim: int,
                 task_configs: Dict, device, adapter_config):
        super().__init__()
        self.encoder_dim = encoder_dim
        self.vilt_encoder = encoder
        self.ordered_cl_tasks = ordered_cl_tasks
        self.task_configs = task_configs
        self.device = device
        self.adapter_config = adapter_config
        self.task_layer_dict = {}
        for task_key in ordered_cl_tasks:
            self.add_task_layer(task_key, task_configs[task_key])
        self.task_layer = nn.ModuleDict(self.task_layer_dict)

    def add_task_layer(self, task_key: str, task_config: Dict):
        """vilt.py:187-219."""
        num_labels = task_config["num_labels"]
        if task_config["model_type"] == "classification":
            num_images = task_config["num_images"]
            clf_layer = nn.Sequential(OrderedDict([
                ("clf_fc0", nn.Linear(self.encoder_dim * num_images, self.encoder_dim * 2)),
                ("clf_norm0", nn.LayerNorm(self.encoder_dim * 2)),
                ("clf_actv0", nn.GELU()),
                ("clf_fc1", nn.Linear(self.encoder_dim * 2, num_labels)),
            ]))
            self.task_layer_dict[task_key] = clf_layer
        elif task_config["model_type"] == "multi-choice":
            clf_layer = nn.Sequential(OrderedDict([
                ("clf_dropout", nn.Dropout(0.1)),
                ("clf_fc0", nn.Linear(self.encoder_dim, 1)),
            ]))
            self.task_layer_dict[task_key] = clf_layer

    def forward(self, task_key: str, images: Optional[List] = None, texts: Optional[List[str]] = None,
                **encodings):
        """vilt.py:221-242 (+ the pre-encoded tensor entry, see module docstring)."""
        task_config = self.task_configs[task_key]
        if task_config["model_type"] != "classification" or task_config["num_images"] != 1:
            raise NotImplementedError("only single-image classification heads are on the DAT VQA path "
                                      "(vilt.py:244-264); multi-choice / multi-image tasks are out of scope")
        return self.forward_single_image(task_key, images, texts, **encodings)

    def forward_single_image(self, task_key: str, images=None, texts=None, **encodings):
        """vilt.py:244-264: (pooler_output, logits)."""
        encoder_output = self.encode(images, texts, **encodings)
        return encoder_output, self.classify(task_key, encoder_output)

    def encode(self, images=None, texts=None, **encodings):
        """The encoder half of ``forward_single_image`` (vilt.py:257-259): pooled output (batch, hidden)."""
        if images is not None:
            encodings = dict(self.vilt_encoder.process_inputs(images, texts))
        return self.vilt_encoder(**encodings)

    def classify(self, task_key: str, encoder_output):
        """The task-head half of ``forward_single_image`` (vilt.py:261-262)."""
        head = self.task_layer[task_key]
        return head(encoder_output.to(head.clf_fc0.weight.dtype))

    def encode_dual(self, **encodings):
        """Pooled outputs of the gating pass (rows [0, batch)) and of the adapter_1 pass (rows [batch,
        2 batch)) from ONE stacked forward: (2 batch, hidden).  Needs the dense fast path (all-ones masks)."""
        if not (encodings.get("dense_masks", False) and self.vilt_encoder.dense_fast_path):
            raise RuntimeError("encode_dual needs pre-encoded batches with all-ones masks (dense_masks=True)")
        for a in self._adapters():
            a.set_dual(True)
        try:
            enc = self.vilt_encoder._dense_forward_stacked(encodings["input_ids"], encodings.get("token_type_ids"),
                                                           encodings["pixel_values"])
        finally:
            for a in self._adapters():
                a.set_dual(False)
        return enc

    def new_step(self, train: bool = False) -> None:
        """Called by the trainer at the start of every train / eval step: drops the per-step embedding cache;
        for a train step also packs the bf16 operands of all 12 sites, both modes, in one launch."""
        from .adapter import refresh_packs
        self.vilt_encoder._embed_cache = None
        if train:
            refresh_packs(self._adapters())

    def end_step(self) -> None:
        """The optimizer has changed the masters: the packed operands are stale from here on."""
        from .adapter import invalidate_packs
        invalidate_packs(self._adapters())

    def gating_forward_is_reusable(self) -> bool:
        """True when the encoder is a deterministic function of (inputs, adapter_0, adapter_2, frozen
        backbone): every ViLT dropout probability is 0 (HF ViltConfig defaults).  The MKD
        schedule's passes A and C then share one encoder forward (TaskTrainer.train_step)."""
        cfg = self.vilt_encoder.vilt.config
        return (float(getattr(cfg, "hidden_dropout_prob", 0.0)) == 0.0
                and float(getattr(cfg, "attention_probs_dropout_prob", 0.0)) == 0.0)

    # ------------------------------------------------------------------ adapter hooks (vilt.py:356-382)
    def add_adapter(self):
        for i in range(len(self.vilt_encoder.vilt.encoder.layer)):
            self.vilt_encoder.vilt.encoder.layer[i].output = Adaptered_ViltOutput(
                self.vilt_encoder.vilt.encoder.layer[i].output, self.adapter_config)

    def _adapters(self):
        return [layer.output.adapter for layer in self.vilt_encoder.vilt.encoder.layer]

    def set_active_adapter(self, name):
        for a in self._adapters():
            a.set_active_adapter(name)

    def activate_gating(self):
        for a in self._adapters():
            a.activate_gating()

    def deactivate_gating(self):
        for a in self._adapters():
            a.deactivate_gating()

    def get_param_adapter(self, name):
        out = []
        for a in self._adapters():
            out.append(getattr(a, f"{name}_down").parameters())
            out.append(getattr(a, f"{name}_up").parameters())
        return out

    # ------------------------------------------------------------------ B200 setup helpers
    def cast_frozen_backbone(self, dtype=torch.bfloat16):
        """bf16 for the frozen ViLT backbone; adapters and task heads keep fp32 masters."""
        for name, p in self.vilt_encoder.vilt.named_parameters():
            if "adapter" not in name:
                p.data = p.data.to(dtype)
        for name, b in self.vilt_encoder.vilt.named_buffers():
            if b.is_floating_point():
                b.data = b.data.to(dtype)
        return self


def load_vilt_encoder(logger, checkpoint_name: str, device, pretrained_vilt_name: str) -> ViltEncoderWrapper:
    """vilt.py:387-418.  ``random`` / ``random-init`` (or any name whose files are not on disk: there is
    no network) builds ViLT-B/32 from ``ViltConfig()`` defaults with seeded random weights."""
    from transformers import ViltConfig, ViltModel
    processor = None
    vilt = None
    if pretrained_vilt_name not in ("random", "random-init"):
        try:
            from transformers import ViltProcessor
            processor = ViltProcessor.from_pretrained(pretrained_vilt_name, local_files_only=True)
            vilt = ViltModel.from_pretrained(pretrained_vilt_name, local_files_only=True)
        except Exception as e:  # noqa: BLE001 - offline box: fall back to the documented random init
            logger.warning("pretrained ViLT %r not available offline (%s): using random-init ViLT-B/32",
                           pretrained_vilt_name, type(e).__name__)
    if vilt is None:
        vilt = ViltModel(ViltConfig())
    encoder = ViltEncoderWrapper(processor, vilt, device)
    if checkpoint_name != pretrained_vilt_name and checkpoint_name not in ("random", "random-init"):
        encoder.load_state_dict(torch.load(checkpoint_name, map_location="cpu"))
    logger.info("Successfully loaded ViLT encoder")
    return encoder


def create_vilt_continual_learner_model(logger, model_name_or_path: str, ordered_cl_tasks: List[str],
                                        model_config: Dict, task_configs: Dict, device):
    """vilt.py:421-452."""
    logger = logger or logging.getLogger(__name__)
    encoder = load_vilt_encoder(logger, checkpoint_name=model_name_or_path, device=device,
                                pretrained_vilt_name=model_name_or_path)
    cl_model = ViltContinualLearner(ordered_cl_tasks=ordered_cl_tasks, encoder=encoder,
                                    encoder_dim=model_config["encoder_dim"], task_configs=task_configs,
                                    device=device,
                                    adapter_config=model_config["adapter_config"] if "adapter_config" in model_config else None)
    logger.info("Successfully created and initialized ViLT Continual Learner model")
    return cl_model


def convert_batch_to_vilt_input_dict(batch: Dict):
    """vilt.py:455-459, plus the pre-encoded form used by the synthetic / pinned-tensor loaders."""
    if "encodings" in batch:
        return dict(batch["encodings"])
    return {"images": batch["images"], "texts": batch["raw_texts"]}
