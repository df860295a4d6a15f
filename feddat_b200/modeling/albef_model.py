"""ALBEF VQA model with the DAT sites (BASELINE configs[2]): ViT-B/16 visual
encoder -> 12-layer BERT question encoder (cross-attention to the image in layers >= fusion_layer) -> 6-layer
BERT answer decoder with the LM head, as wired by reference src/modeling/models/albef_model.py:13-156, built
from the reference's vendored vit.py / xbert.py STRUCTURE so that every state-dict key matches:

    visual_encoder.{patch_embed.proj, cls_token, pos_embed, blocks.{i}.{norm1, attn.qkv, attn.proj, norm2,
                    mlp.fc1, mlp.fc2, adapter.adapter_{j}_{down,up}}, norm}                       (vit.py:112-190)
    text_encoder.{embeddings.*, encoder.layer.{i}.{attention.{self.{query,key,value}, output.{dense,LayerNorm}},
                  crossattention.* (i >= fusion_layer), intermediate.dense,
                  output.{dense, LayerNorm, adapter.*}}}                                            (xbert.py:170-635)
    text_decoder.{bert.<same>, cls.predictions.{transform.{dense,LayerNorm}, decoder, bias}}      (xbert.py:670-697,1187)

(ALBEF checkpoints and the round loop's substring selection -- 'adapter_1' communicated, '.cls.' personal,
main.py:127-128,154-163 -- apply unchanged; tests/test_albef_gpu.py pins the whole forward / train step against
the reference modules filled with the same by-name weights.)

What is NOT restated: the tokenizer (host text preprocessing: the forward takes token ids), the momentum-distilled
variant (``distill=True``; FedDAT's ALBEF runs use albef_no_distill, src/train_albef.sh:3), head pruning, relative
position embeddings, attention-map hooks.  Attention is SDPA over additive masks (same arithmetic as
softmax(QK^T / sqrt d + mask) V); the adapter sites run the sm_100a DAT kernels.
"""
from __future__ import annotations

import math
from functools import partial
from types import SimpleNamespace
from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.attention import SDPBackend, sdpa_kernel

from .albef_sites import BertOutput, Block
from .fused_ln import layer_norm_of_sum

FUSE_PROJECTIONS = True      # frozen q | k | v (self-attention) and k | v (cross-attention) projections as one GEMM
# torch SDPA backend of the BERT-side attention calls with <= 128 queries and keys (None = torch's own choice)
SMALL_ATTENTION_BACKEND = SDPBackend.EFFICIENT_ATTENTION

# reference src/configs/model_configs.py:40-60
CONFIG_BERT = {
    "attention_probs_dropout_prob": 0.1, "hidden_act": "gelu", "hidden_dropout_prob": 0.1, "hidden_size": 768,
    "initializer_range": 0.02, "intermediate_size": 3072, "layer_norm_eps": 1e-12, "max_position_embeddings": 512,
    "num_attention_heads": 12, "num_hidden_layers": 12, "pad_token_id": 0, "type_vocab_size": 2, "vocab_size": 30522,
    "fusion_layer": 6, "encoder_width": 768,
}
PAD_TOKEN_ID, SEP_TOKEN_ID, CLS_TOKEN_ID = 0, 102, 101        # bert-base-uncased


def bert_config(**over) -> SimpleNamespace:
    cfg = dict(CONFIG_BERT)
    cfg.update(over)
    return SimpleNamespace(**cfg)


# ------------------------------------------------------------------------------------------ visual encoder
class PatchEmbed(nn.Module):
    """timm PatchEmbed as vit.py:144-146 uses it: Conv2d(in, dim, patch, patch) + flatten."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        self.img_size, self.patch_size = (img_size, img_size), (patch_size, patch_size)
        self.num_patches = (img_size // patch_size) ** 2
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        # non-overlapping patches: the convolution is ONE [B * patches, 3 * p * p] x [3 * p * p, dim] GEMM
        b, c, hh, ww = x.shape
        p = self.patch_size[0]
        h, w = hh // p, ww // p
        patches = x.to(self.proj.weight.dtype).reshape(b, c, h, p, w, p).permute(0, 2, 4, 1, 3, 5).reshape(b * h * w, c * p * p)
        return F.linear(patches, self.proj.weight.view(self.proj.out_channels, -1), self.proj.bias).view(b, h * w, -1)


class VisionTransformer(nn.Module):
    """vit.py:112-190."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0,
                 qkv_bias=True, qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0, norm_layer=None, adapter_config=None):
        super().__init__()
        self.num_features = self.embed_dim = embed_dim
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + 1, embed_dim))
        self.pos_drop = nn.Dropout(p=drop_rate)
        self.blocks = nn.ModuleList([
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                  drop=drop_rate, attn_drop=attn_drop_rate, norm_layer=norm_layer, adapter_config=adapter_config)
            for _ in range(depth)])
        self.norm = norm_layer(embed_dim)
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.trunc_normal_(self.cls_token, std=0.02)
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m):                                             # vit.py:161-168
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def forward(self, x):                                             # vit.py:174-189
        b = x.shape[0]
        x = self.patch_embed(x)
        x = torch.cat((self.cls_token.expand(b, -1, -1).to(x.dtype), x), dim=1)
        x = self.pos_drop(x + self.pos_embed[:, :x.size(1), :].to(x.dtype))
        for blk in self.blocks:
            x = blk(x)
        return self.norm(x)


# ------------------------------------------------------------------------------------------ BERT with cross-attention
class BertEmbeddings(nn.Module):
    """xbert.py:170-216."""

    def __init__(self, config):
        super().__init__()
        self.word_embeddings = nn.Embedding(config.vocab_size, config.hidden_size, padding_idx=config.pad_token_id)
        self.position_embeddings = nn.Embedding(config.max_position_embeddings, config.hidden_size)
        self.token_type_embeddings = nn.Embedding(config.type_vocab_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.register_buffer("position_ids", torch.arange(config.max_position_embeddings).expand((1, -1)))

    def forward(self, input_ids):
        seq = input_ids.shape[1]
        e = self.word_embeddings(input_ids) + self.token_type_embeddings(torch.zeros_like(input_ids))
        e = e + self.position_embeddings(self.position_ids[:, :seq])
        return self.dropout(self.LayerNorm(e))


class BertSelfAttention(nn.Module):
    """xbert.py:219-347: softmax(Q K^T / sqrt(d) + mask) V; keys / values from the encoder states when the module
    is a cross-attention."""

    def __init__(self, config, is_cross_attention):
        super().__init__()
        self.num_attention_heads = config.num_attention_heads
        self.attention_head_size = config.hidden_size // config.num_attention_heads
        kv_in = config.encoder_width if is_cross_attention else config.hidden_size
        self.query = nn.Linear(config.hidden_size, config.hidden_size)
        self.key = nn.Linear(kv_in, config.hidden_size)
        self.value = nn.Linear(kv_in, config.hidden_size)
        self.dropout = nn.Dropout(config.attention_probs_dropout_prob)

    def _heads(self, x):
        b, s, _ = x.shape
        return x.view(b, s, self.num_attention_heads, self.attention_head_size).transpose(1, 2)

    def _cat(self, mods, tag):
        """Concatenated weight / bias of FROZEN projections that read the same tensor (q | k | v of a self-attention,
        k | v of a cross-attention): one GEMM instead of two or three, and in backward one GEMM over the stacked
        gradient instead of one per projection plus the adds.  Cached on the module; rebuilt when a tensor changes."""
        ps = [m.weight for m in mods] + [m.bias for m in mods]
        key = tuple((p.data_ptr(), p._version, p.dtype) for p in ps)
        cached = getattr(self, "_feddat_cat_" + tag, None)
        if cached is None or cached[0] != key:
            with torch.no_grad():
                n = len(mods)
                cached = (key, torch.cat([p.detach() for p in ps[:n]], 0).contiguous(),
                          torch.cat([p.detach() for p in ps[n:]], 0).contiguous())
            setattr(self, "_feddat_cat_" + tag, cached)
        return cached[1], cached[2]

    def forward(self, hidden_states, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None):
        kv = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        mask = attention_mask if encoder_hidden_states is None else encoder_attention_mask
        heads, hd = self.num_attention_heads, self.attention_head_size
        frozen = FUSE_PROJECTIONS and not any(p.requires_grad for m in (self.query, self.key, self.value) for p in m.parameters())
        if frozen and encoder_hidden_states is None:
            w, b = self._cat((self.query, self.key, self.value), "qkv")
            bs, sq, _ = hidden_states.shape
            q, k, v = (t.transpose(1, 2) for t in F.linear(hidden_states, w, b).view(bs, sq, 3, heads, hd).unbind(2))
        elif frozen:
            w, b = self._cat((self.key, self.value), "kv")
            kv = kv.to(hidden_states.dtype)
            bs, sk, _ = kv.shape
            q = self._heads(self.query(hidden_states))
            k, v = (t.transpose(1, 2) for t in F.linear(kv, w, b).view(bs, sk, 2, heads, hd).unbind(2))
        else:
            q, k, v = self._heads(self.query(hidden_states)), self._heads(self.key(kv.to(hidden_states.dtype))), \
                self._heads(self.value(kv.to(hidden_states.dtype)))
        if mask is not None:
            mask = mask.to(q.dtype)
        p_drop = self.dropout.p if self.training else 0.0
        if SMALL_ATTENTION_BACKEND is not None and q.is_cuda and q.shape[2] <= 128 and k.shape[2] <= 128:
            # a few dozen tokens per sequence (questions <= 25, answers ~6): cuDNN's flash kernels with a mask and
            # dropout take 47-64 us per call here, the memory-efficient backend 15-17 us (fwd + bwd: 80-107 -> 34-37;
            # scripts/sdpa_backends_albef.py) -- 90 forward and 60 backward calls per train step
            with sdpa_kernel(SMALL_ATTENTION_BACKEND):
                ctx = F.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=p_drop)
        else:
            ctx = F.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=p_drop)
        b, _, s, _ = ctx.shape
        return ctx.transpose(1, 2).reshape(b, s, -1)


class BertSelfOutput(nn.Module):
    """xbert.py:350-361."""

    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, hidden_states, input_tensor):
        return layer_norm_of_sum(self.LayerNorm, self.dropout(self.dense(hidden_states)), input_tensor)


class BertAttention(nn.Module):
    """xbert.py:364-411."""

    def __init__(self, config, is_cross_attention=False):
        super().__init__()
        self.self = BertSelfAttention(config, is_cross_attention)
        self.output = BertSelfOutput(config)

    def forward(self, hidden_states, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None):
        return self.output(self.self(hidden_states, attention_mask, encoder_hidden_states, encoder_attention_mask),
                           hidden_states)


class BertIntermediate(nn.Module):
    """xbert.py:414-426 (hidden_act 'gelu' = the exact erf form)."""

    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.intermediate_size)
        if config.hidden_act != "gelu":
            raise NotImplementedError("ALBEF's BERT uses hidden_act='gelu' (model_configs.py:45)")

    def forward(self, hidden_states):
        return F.gelu(self.dense(hidden_states))


class BertLayer(nn.Module):
    """xbert.py:448-525: self-attention, cross-attention in layers >= fusion_layer, feed-forward whose output block
    carries the DAT site (albef_sites.BertOutput, xbert.py:428-445)."""

    def __init__(self, config, layer_num):
        super().__init__()
        self.attention = BertAttention(config)
        self.has_cross_attention = layer_num >= config.fusion_layer
        if self.has_cross_attention:
            self.crossattention = BertAttention(config, is_cross_attention=True)
        self.intermediate = BertIntermediate(config)
        self.output = BertOutput(config)

    def forward(self, hidden_states, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None):
        attention_output = self.attention(hidden_states, attention_mask)
        if self.has_cross_attention:
            assert encoder_hidden_states is not None, "encoder_hidden_states must be given for cross-attention layers"
            attention_output = self.crossattention(attention_output, attention_mask, encoder_hidden_states,
                                                   encoder_attention_mask)
        return self.output(self.intermediate(attention_output), attention_output)


class BertEncoder(nn.Module):
    """xbert.py:528-635."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.layer = nn.ModuleList([BertLayer(config, i) for i in range(config.num_hidden_layers)])

    def forward(self, hidden_states, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None,
                mode="multi_modal"):
        first, last = {"text": (0, self.config.fusion_layer), "fusion": (self.config.fusion_layer, len(self.layer)),
                       "multi_modal": (0, len(self.layer))}[mode]
        for i in range(first, last):
            hidden_states = self.layer[i](hidden_states, attention_mask, encoder_hidden_states, encoder_attention_mask)
        return hidden_states


def _init_bert_weights(module, std):
    """xbert.py BertPreTrainedModel._init_weights."""
    if isinstance(module, (nn.Linear, nn.Embedding)):
        module.weight.data.normal_(mean=0.0, std=std)
    elif isinstance(module, nn.LayerNorm):
        module.bias.data.zero_()
        module.weight.data.fill_(1.0)
    if isinstance(module, nn.Linear) and module.bias is not None:
        module.bias.data.zero_()


class BertModel(nn.Module):
    """xbert.py:803-1083 (no pooling layer on the ALBEF path: albef_model.py:42)."""

    def __init__(self, config, add_pooling_layer=False):
        super().__init__()
        if add_pooling_layer:
            raise NotImplementedError("ALBEF builds its BERTs with add_pooling_layer=False")
        self.config = config
        self.embeddings = BertEmbeddings(config)
        self.encoder = BertEncoder(config)
        self.apply(partial(_init_bert_weights, std=config.initializer_range))

    @staticmethod
    def extended_attention_mask(attention_mask, is_decoder, dtype):
        """xbert.py:880-938: [B, S] padding mask -> additive [B, 1, S | 1, S] mask, causal for a decoder."""
        if is_decoder:
            b, s = attention_mask.shape
            ids = torch.arange(s, device=attention_mask.device)
            causal = (ids[None, None, :].repeat(b, s, 1) <= ids[None, :, None]).to(attention_mask.dtype)
            ext = causal[:, None, :, :] * attention_mask[:, None, None, :]
        else:
            ext = attention_mask[:, None, None, :]
        return (1.0 - ext.to(dtype)) * -10000.0

    def forward(self, input_ids, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None,
                is_decoder=False, mode="multi_modal"):
        if attention_mask is None:
            attention_mask = torch.ones_like(input_ids)
        x = self.embeddings(input_ids)
        ext = self.extended_attention_mask(attention_mask, is_decoder, x.dtype)
        enc_ext = None
        if encoder_hidden_states is not None and encoder_attention_mask is not None:
            # transformers' invert_attention_mask: a large negative number on masked keys.  No mask given = every
            # encoder position visible (the reference builds an all-ones mask, xbert.py:1030-1040, whose additive form
            # is all zeros): nothing is added, and the attention kernel runs without a mask operand
            enc_ext = (1.0 - encoder_attention_mask[:, None, None, :].to(x.dtype)) * -10000.0
        return self.encoder(x, ext, encoder_hidden_states, enc_ext, mode=mode)


class BertPredictionHeadTransform(nn.Module):
    """xbert.py:655-668."""

    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)

    def forward(self, hidden_states):
        return self.LayerNorm(F.gelu(self.dense(hidden_states)))


class BertLMPredictionHead(nn.Module):
    """xbert.py:670-689: the output-only bias is registered twice (``bias`` and ``decoder.bias``), as upstream."""

    def __init__(self, config):
        super().__init__()
        self.transform = BertPredictionHeadTransform(config)
        self.decoder = nn.Linear(config.hidden_size, config.vocab_size, bias=False)
        self.bias = nn.Parameter(torch.zeros(config.vocab_size))
        self.decoder.bias = self.bias

    def forward(self, hidden_states):
        return self.decoder(self.transform(hidden_states))


class BertOnlyMLMHead(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.predictions = BertLMPredictionHead(config)

    def forward(self, sequence_output):
        return self.predictions(sequence_output)


class BertLMHeadModel(nn.Module):
    """xbert.py:1187-1315.  ``forward`` returns the prediction scores [n_seq, La, vocab]; the next-token loss of
    xbert.py:1287-1297 lives in ``lm_loss`` (torch) and, on the training path, in the fused KL + CE kernel."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.bert = BertModel(config, add_pooling_layer=False)
        self.cls = BertOnlyMLMHead(config)
        self.apply(partial(_init_bert_weights, std=config.initializer_range))

    def forward(self, input_ids, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None,
                is_decoder=True, mode="multi_modal"):
        h = self.bert(input_ids, attention_mask, encoder_hidden_states, encoder_attention_mask, is_decoder=is_decoder,
                      mode=mode)
        if h.is_cuda and h.dtype == torch.bfloat16 and self.cls.predictions.decoder.weight.dtype == torch.float32:
            # the LM head is TRAINABLE on the FedDAT path ('.cls.' keys, main.py:127-128,248-250): fp32 masters, bf16
            # arithmetic -- what accelerate's mixed precision does to the whole reference model
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return self.cls(h)
        return self.cls(h)

    @staticmethod
    def lm_loss(prediction_scores, labels):
        """xbert.py:1287-1295 with reduction='none': per-sequence sum of the shifted token cross-entropies."""
        shifted = prediction_scores[:, :-1, :].float()
        tgt = labels[:, 1:]
        loss = F.cross_entropy(shifted.reshape(-1, shifted.shape[-1]), tgt.reshape(-1), reduction="none")
        return loss.view(prediction_scores.size(0), -1).sum(1)


# ------------------------------------------------------------------------------------------ ALBEF
class LazyAnswerLoss:
    """What ``ALBEF.forward(train=True, defer_loss=True)`` returns in place of the scalar loss: everything the
    fused MKD head (feddat_mkd_ce_loss: KL + weighted token CE + d/dlogits in one pass) needs.  ``value()`` is the
    reference's own expression (albef_model.py:142-143) in torch."""

    def __init__(self, prediction_scores, labels, weights, batch_size):
        self.prediction_scores, self.labels, self.weights, self.batch_size = prediction_scores, labels, weights, batch_size

    def value(self):
        per_seq = BertLMHeadModel.lm_loss(self.prediction_scores, self.labels)
        return (self.weights * per_seq).sum() / self.batch_size


class ALBEF(nn.Module):
    """albef_model.py:13-228 (distill=False)."""

    def __init__(self, config: Dict, tokenizer=None):
        super().__init__()
        if config.get("distill", False):
            raise NotImplementedError("the momentum-distilled ALBEF (albef_distill) is not on the FedDAT path "
                                      "(src/train_albef.sh uses albef_no_distill)")
        self.tokenizer = tokenizer
        self.pad_token_id = getattr(tokenizer, "pad_token_id", PAD_TOKEN_ID)
        self.distill = False
        adapter_config = config.get("adapter_config")
        self.visual_encoder = VisionTransformer(
            img_size=config["image_res"], patch_size=16, embed_dim=768, depth=config.get("vit_depth", 12), num_heads=12,
            mlp_ratio=4, qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6), adapter_config=adapter_config)
        enc = bert_config(**config["bert_config"])
        dec = bert_config(**config["bert_config"])
        dec.fusion_layer = 0                                          # albef_model.py:32-33
        dec.num_hidden_layers = config.get("decoder_layers", 6)
        if adapter_config is not None:
            enc.adapter_config = adapter_config
            dec.adapter_config = adapter_config
        self.text_encoder = BertModel(enc, add_pooling_layer=False)
        self.text_decoder = BertLMHeadModel(dec)

    def forward(self, image, question, answer=None, alpha=0, k=None, weights=None, train=True, defer_loss=False,
                answer_index=None, image_embeds=None):
        """``question`` / ``answer``: objects with ``input_ids`` and ``attention_mask`` (a tokenizer's BatchEncoding or
        a namespace of tensors).  train: (loss, logits[:, :-1]); eval: rank_answer's (topk_ids, topk_probs).
        ``image_embeds``: the visual encoder's output for ``image`` when the caller already has it (the MKD schedule's
        passes A and C see the same image through the same gating adapters: TaskTrainer runs the ViT once for both)."""
        if image_embeds is None:
            image_embeds = self.visual_encoder(image)
        # albef_model.py:95-96 builds an all-ones image mask; "no mask" is the same attention (see BertModel.forward)
        image_atts = None
        question_states = self.text_encoder(question.input_ids, attention_mask=question.attention_mask,
                                            encoder_hidden_states=image_embeds, encoder_attention_mask=image_atts)
        if not train:
            return self.rank_answer(question_states, question.attention_mask, answer.input_ids, answer.attention_mask, k)
        answer_targets = answer.input_ids.masked_fill(answer.input_ids == self.pad_token_id, -100)
        # one copy of the question states per answer of that question (albef_model.py:92-98)
        # (``answer_index`` = that map as a tensor made with the batch: repeat_interleave over a device tensor of
        # counts has a data-dependent output size, i.e. a host sync per forward and no CUDA-graph capture)
        idx = answer_index if answer_index is not None else torch.repeat_interleave(
            torch.arange(len(k), device=image.device), torch.as_tensor(k, device=image.device))
        scores = self.text_decoder(answer.input_ids, attention_mask=answer.attention_mask,
                                   encoder_hidden_states=question_states[idx],
                                   encoder_attention_mask=question.attention_mask[idx])
        lazy = LazyAnswerLoss(scores, answer_targets, weights, image.size(0))
        # reference returns logits[:, :-1, :].contiguous(); the view avoids a copy of the largest tensor of the model
        return (lazy if defer_loss else lazy.value()), scores[:, :-1, :]

    @torch.no_grad()
    def rank_answer(self, question_states, question_atts, answer_ids, answer_atts, k):
        """albef_model.py:171-228: first-token top-k over the answer list, then re-rank by sequence likelihood."""
        num_ques = question_states.size(0)
        start_ids = answer_ids[0, 0].repeat(num_ques, 1)              # bos token
        logits = self.text_decoder(start_ids, encoder_hidden_states=question_states,
                                   encoder_attention_mask=question_atts)[:, 0, :]
        prob_first = F.softmax(logits.float(), dim=1).index_select(dim=1, index=answer_ids[:, 1])
        topk_probs, topk_ids = prob_first.topk(k, dim=1)
        input_ids = answer_ids.index_select(0, topk_ids.reshape(-1))
        input_atts = answer_atts.index_select(0, topk_ids.reshape(-1))
        targets = input_ids.masked_fill(input_ids == self.pad_token_id, -100)
        rep = torch.arange(num_ques, device=question_states.device).repeat_interleave(k)      # tile(x, 0, k)
        scores = self.text_decoder(input_ids, attention_mask=input_atts, encoder_hidden_states=question_states[rep],
                                   encoder_attention_mask=question_atts[rep])
        answer_loss = BertLMHeadModel.lm_loss(scores, targets).view(input_ids.size(0), -1)
        log_probs = torch.cat([topk_probs.view(-1, 1).log(), -answer_loss], dim=1).sum(1).view(num_ques, k)
        topk_probs, rerank = F.softmax(log_probs, dim=-1).topk(k, dim=1)
        return torch.gather(topk_ids, 1, rerank), topk_probs
