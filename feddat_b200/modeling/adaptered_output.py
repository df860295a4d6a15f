"""Adapter injection wrappers (mirror of reference src/modeling/adaptered_output.py:55-79)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .adapter import Adapter


class Adaptered_ViltOutput(nn.Module):
    """Wraps HF ``ViltOutput`` (``dense`` 3072->768 + ``dropout``): dense, dropout, + residual, then the
    DAT adapter with the block output as both hidden state and residual
    (adaptered_output.py:73-79).  The wrapped layer moves to ``.layer`` so state-dict keys become
    ``...output.layer.dense.*`` / ``...output.adapter.adapter_{j}_{down,up}.*`` exactly as upstream."""

    def __init__(self, layer, adapter_config) -> None:
        super().__init__()
        self.layer = layer
        self.adapter = Adapter(**adapter_config, model_dim=768)

    def forward(self, hidden_states: torch.Tensor, input_tensor: torch.Tensor) -> torch.Tensor:
        hidden_states = self.layer.dense(hidden_states)
        hidden_states = self.layer.dropout(hidden_states)
        hidden_states = hidden_states + input_tensor
        hidden_states = self.adapter(hidden_states, hidden_states)
        return hidden_states


class Adaptered_BertOutput(nn.Module):
    """BERT-style output block: dense, dropout, then LN / adapter / LN through
    ``Adapter.adapter_layer_forward_bert`` (adaptered_output.py:55-65; the same call ALBEF's
    xbert.BertOutput makes at xbert.py:438-445)."""

    def __init__(self, layer, adapter_config):
        super().__init__()
        self.layer = layer
        self.adapter = Adapter(**adapter_config, model_dim=768)

    def forward(self, hidden_states, input_tensor):
        hidden_states = self.layer.dense(hidden_states)
        hidden_states = self.layer.dropout(hidden_states)
        hidden_states = self.adapter.adapter_layer_forward_bert(hidden_states, input_tensor, self.layer.LayerNorm)
        return hidden_states
