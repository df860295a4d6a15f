"""ALBEF adapter-injection sites, with the DAT operator running as the
sm_100a kernels behind ``Adapter``.

* ``Block``       -- the ViT encoder block of ALBEF's visual encoder with the adapter after the MLP
                     residual (reference src/modeling/models/vit.py:78-110): same constructor keywords
                     (``adapter_config=None`` keeps the plain block), same sub-module names
                     (``norm1, attn.qkv, attn.proj, norm2, mlp.fc1, mlp.fc2, adapter.*``) so ALBEF
                     checkpoints and the round loop's substring selection
                     (``...visual_encoder.blocks.{i}.adapter.adapter_{j}_{down,up}``) apply unchanged.
* ``BertOutput``  -- the feed-forward output block of ALBEF's text encoder / decoder (reference
                     src/modeling/models/xbert.py:428-445): ``config.adapter_config`` switches on the
                     double-LayerNorm adapter wrapper ``Adapter.adapter_layer_forward_bert``
                     (adapter.py:97-116; the residual handed to the operator is the FFN output, not
                     the operator's input).

The attention / MLP / LayerNorm parts are frozen backbone and stay PyTorch ops (SDPA for the
softmax(QK^T)V product); the full ALBEF model wrapper (tokenizer, cross-attention fusion, LM head) is
not on the DAT hot path and is not rebuilt here.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .adapter import Adapter


class Mlp(nn.Module):
    """vit.py:12-31."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        return self.drop(self.fc2(self.drop(self.act(self.fc1(x)))))


class Attention(nn.Module):
    """vit.py:34-75: softmax(q k^T * scale) v through SDPA (frozen backbone op)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        self.num_heads = num_heads
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x, register_hook=False):
        b, n, c = x.shape
        qkv = self.qkv(x).reshape(b, n, 3, self.num_heads, c // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        if register_hook:
            raise NotImplementedError("register_hook (Grad-CAM attention maps, vit.py:48-58,70-72) is a visualisation "
                                      "aid outside the FedDAT training path")
        x = F.scaled_dot_product_attention(q, k, v, dropout_p=self.attn_drop.p if self.training else 0.0,
                                           scale=self.scale)
        x = x.transpose(1, 2).reshape(b, n, c)
        return self.proj_drop(self.proj(x))


class Block(nn.Module):
    """vit.py:78-110."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop=0.0, attn_drop=0.0,
                 drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm, adapter_config=None):
        super().__init__()
        if drop_path > 0.0:
            raise NotImplementedError("stochastic depth is not used on the FedDAT path (albef_model.py:24-28)")
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale,
                              attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        if adapter_config is None:
            self.adaptered = False
        else:
            self.adaptered = True
            self.adapter = Adapter(**adapter_config, model_dim=dim)

    def forward(self, x, register_hook=False):
        x = x + self.attn(self.norm1(x), register_hook=register_hook)
        x = x + self.mlp(self.norm2(x))
        if self.adaptered:
            x = self.adapter(x, x)                               # vit.py:107: residual IS the input
        return x


class BertOutput(nn.Module):
    """xbert.py:428-445."""

    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.intermediate_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        if hasattr(config, "adapter_config"):
            self.adapter = Adapter(**config.adapter_config, model_dim=config.hidden_size)

    def forward(self, hidden_states, input_tensor):
        hidden_states = self.dropout(self.dense(hidden_states))
        if hasattr(self, "adapter"):
            return self.adapter.adapter_layer_forward_bert(hidden_states, input_tensor, self.LayerNorm)
        return self.LayerNorm(hidden_states + input_tensor)


class AdapterHooks:
    """Learner-level adapter hooks (reference src/modeling/albef.py:139-167, src/modeling/vilt.py:363-382) for
    ANY module tree containing ``Adapter`` sites -- ALBEF's 12 ViT blocks + 12 text-encoder + 6 text-decoder
    output layers, or ViLT's 12 -- found by walking the sub-modules instead of hard-coded attribute paths.
    Mix into (or wrap) the learner: ``hooks = AdapterHooks(model)``."""

    def __init__(self, root: nn.Module):
        self._root = root

    def adapters(self):
        return [m for m in self._root.modules() if isinstance(m, Adapter)]

    def set_active_adapter(self, name):
        for a in self.adapters():
            a.set_active_adapter(name)

    def activate_gating(self):
        for a in self.adapters():
            a.activate_gating()

    def deactivate_gating(self):
        for a in self.adapters():
            a.deactivate_gating()

    def get_param_adapter(self, name):
        out = []
        for a in self.adapters():
            out.append(getattr(a, f"{name}_down").parameters())
            out.append(getattr(a, f"{name}_up").parameters())
        return out
