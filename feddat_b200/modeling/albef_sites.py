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

The frozen block around the ViT site runs on this repo's kernels when its preconditions hold (bf16 CUDA activations,
frozen bf16 parameters, no dropout -- ``Block._fast_forward``): LayerNorm-before and "first residual + LayerNorm-after"
as one launch each (feddat_ln_fwd / feddat_ln_bwd), the 768 -> 3072 GEMM with the exact GELU in its epilogue and the
backward's GELU' in its first GEMM (feddat_mlp_fc1_gelu_fwd / feddat_mlp_fc2_dgelu_bwd), "fc2 + bias + second
residual" as one GEMM.  The 577-token softmax(QK^T)V stays torch SDPA (cuDNN; the own attention kernels cover
sequences up to 256 keys); stock PyTorch modules otherwise.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .adapter import Adapter

FAST_BLOCK = True      # A/B switch (tests, bench): False = the stock PyTorch modules of the ViT block


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
        # unbind + transpose (views): autograd's backward is ONE stack into the [b, n, 3, heads, hd] layout the qkv
        # GEMM's data gradient reads -- indexing qkv[0], qkv[1], qkv[2] costs three zero-fills, three slice copies and
        # two full-size adds per block and backward pass
        q, k, v = (t.transpose(1, 2) for t in self.qkv(x).view(b, n, 3, self.num_heads, c // self.num_heads).unbind(2))
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
        if FAST_BLOCK and not register_hook and self._fast_ok(x):
            return self._fast_forward(x)
        x = x + self.attn(self.norm1(x), register_hook=register_hook)
        x = x + self.mlp(self.norm2(x))
        if self.adaptered:
            x = self.adapter(x, x)                               # vit.py:107: residual IS the input
        return x

    def _fast_ok(self, x) -> bool:
        from . import fused_ln
        fc1, fc2 = self.mlp.fc1, self.mlp.fc2
        frozen = not any(p.requires_grad for m in (self.norm1, self.norm2, fc1, fc2) for p in m.parameters())
        no_drop = not self.training or (self.mlp.drop.p == 0.0 and self.attn.proj_drop.p == 0.0)
        return (frozen and no_drop and isinstance(self.norm1, nn.LayerNorm) and isinstance(self.norm2, nn.LayerNorm)
                and fused_ln._usable(self.norm1, x) and fused_ln._usable(self.norm2, x)
                and isinstance(self.mlp.act, nn.GELU) and self.mlp.act.approximate == "none"
                and fc1.bias is not None and fc2.bias is not None and fc1.weight.dtype == torch.bfloat16
                and fc2.weight.dtype == torch.bfloat16 and fc1.weight.is_contiguous() and fc2.weight.is_contiguous()
                and fc1.out_features % 256 == 0 and fc1.in_features % 64 == 0 and fc2.in_features == fc1.out_features)

    def _fast_forward(self, x):
        """The same block (vit.py:99-110) on the fused kernels of modeling/fused_ln.py -- see the module docstring."""
        from . import fused_ln
        fc1, fc2 = self.mlp.fc1, self.mlp.fc2
        ln1, x = fused_ln.layer_norm_pass(self.norm1, x)
        attn = self.attn(ln1)
        res_b, ln2 = fused_ln.add_layer_norm(self.norm2, attn, x, bias2=fc2.bias)
        h = fused_ln._FrozenMlp.apply(ln2, res_b, fc1.weight, fused_ln._bias_f32(fc1), fc2.weight,
                                      fused_ln._transposed_weight(fc2))
        if self.adaptered:
            h = self.adapter(h, h)                               # vit.py:107: residual IS the input
        return h


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
        from .fused_ln import layer_norm_of_sum
        return layer_norm_of_sum(self.LayerNorm, hidden_states, input_tensor)


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
