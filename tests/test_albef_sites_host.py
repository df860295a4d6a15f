"""Host-side contract of the ALBEF injection sites (CPU): sub-module / state-dict names of the
reference (vit.py:78-110, xbert.py:428-445), the un-adaptered paths, and the oracle's BERT-site wrapper
pinned to the golden of the reference's own BertOutput."""
import types
from functools import partial
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

import oracle
from feddat_b200.modeling.albef_sites import BertOutput, Block
from tests.golden_inputs import albef_site_inputs, fill_params

NAMES = ["adapter_0", "adapter_1", "adapter_2"]
GOLD = np.load(Path(__file__).resolve().parent / "golden" / "albef_site_golden.npz")


def test_block_keys_and_plain_path():
    blk = Block(dim=768, num_heads=12, mlp_ratio=4, qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6),
                adapter_config={"names": NAMES, "device": "cpu", "rank": 64})
    keys = set(blk.state_dict().keys())
    for k in ("norm1.weight", "attn.qkv.weight", "attn.qkv.bias", "attn.proj.weight", "norm2.bias",
              "mlp.fc1.weight", "mlp.fc2.bias", "adapter.adapter_1_down.weight", "adapter.adapter_2_up.bias"):
        assert k in keys, k
    assert blk.adaptered and not blk.adapter.adapter_2_down.weight.requires_grad
    plain = Block(dim=768, num_heads=12, mlp_ratio=4, qkv_bias=True)
    assert not plain.adaptered and not hasattr(plain, "adapter")
    x = torch.randn(2, 5, 768)
    y_sdpa = plain(x)
    # SDPA == the explicit softmax(q k^T * scale) v of vit.py:65-73
    a = plain.attn
    qkv = a.qkv(plain.norm1(x)).reshape(2, 5, 3, 12, 64).permute(2, 0, 3, 1, 4)
    att = ((qkv[0] @ qkv[1].transpose(-2, -1)) * a.scale).softmax(dim=-1)
    h = x + a.proj((att @ qkv[2]).transpose(1, 2).reshape(2, 5, 768))
    assert torch.allclose(y_sdpa, h + plain.mlp(plain.norm2(h)), atol=1e-5)


def test_bert_output_keys_and_plain_path():
    cfg = types.SimpleNamespace(intermediate_size=3072, hidden_size=768, layer_norm_eps=1e-12, hidden_dropout_prob=0.0)
    plain = BertOutput(cfg)
    assert not hasattr(plain, "adapter")
    h, x = torch.randn(2, 3, 3072), torch.randn(2, 3, 768)
    assert torch.allclose(plain(h, x), plain.LayerNorm(plain.dense(h) + x))
    cfg.adapter_config = {"names": NAMES, "device": "cpu", "rank": 64}
    ad = BertOutput(cfg)
    assert {"dense.weight", "LayerNorm.weight", "adapter.adapter_0_up.weight"} <= set(ad.state_dict().keys())


def test_oracle_bert_site_matches_reference_bert_output():
    """oracle.adapter_layer_forward_bert fed with the dense output == the reference's BertOutput."""
    cfg = types.SimpleNamespace(intermediate_size=3072, hidden_size=768, layer_norm_eps=1e-12, hidden_dropout_prob=0.0,
                                adapter_config={"names": NAMES, "device": "cpu", "rank": int(GOLD["meta_rank"][0])})
    bout = BertOutput(cfg)
    fill_params(bout, seed=12)
    inp = albef_site_inputs()
    sd = {k: v.numpy() for k, v in bout.state_dict().items()}
    ffn = inp["bert_h"].astype(np.float64) @ sd["dense.weight"].T.astype(np.float64) + sd["dense.bias"]

    def br(n):
        return (sd[f"adapter.{n}_down.weight"], sd[f"adapter.{n}_down.bias"], sd[f"adapter.{n}_up.weight"],
                sd[f"adapter.{n}_up.bias"])

    for mode, branches, gating in (("single_adapter_1", [br("adapter_1")], False),
                                   ("gating", [br("adapter_0"), br("adapter_2")], True)):
        got = oracle.adapter_layer_forward_bert(ffn, inp["bert_x"], sd["LayerNorm.weight"], sd["LayerNorm.bias"],
                                                1e-12, branches, gating)
        want = GOLD[f"bert_output/{mode}/out"]
        assert np.abs(got - want).max() / np.abs(want).max() < 1e-4


def test_adapter_hooks_walk_every_site():
    """AdapterHooks (albef.py:139-167): mode switches reach every Adapter of a mixed ViT / BERT module tree."""
    from feddat_b200.modeling.albef_sites import AdapterHooks
    acfg = {"names": NAMES, "device": "cpu", "rank": 16}
    cfg = types.SimpleNamespace(intermediate_size=3072, hidden_size=768, layer_norm_eps=1e-12, hidden_dropout_prob=0.0,
                                adapter_config=acfg)
    tree = nn.ModuleDict({
        "visual_encoder": nn.ModuleList([Block(dim=768, num_heads=12, qkv_bias=True, adapter_config=acfg) for _ in range(2)]),
        "text_encoder": nn.ModuleList([BertOutput(cfg) for _ in range(2)]),
        "text_decoder": nn.ModuleList([BertOutput(cfg)]),
    })
    hooks = AdapterHooks(tree)
    ads = hooks.adapters()
    assert len(ads) == 5
    hooks.activate_gating(); hooks.set_active_adapter("adapter_0")
    assert all(a.gating and a.adapter_0_up.weight.requires_grad and not a.adapter_1_up.weight.requires_grad for a in ads)
    hooks.deactivate_gating(); hooks.set_active_adapter("adapter_1")
    assert all((not a.gating) and a._active_name == "adapter_1" and not a.adapter_0_down.bias.requires_grad for a in ads)
    assert len(hooks.get_param_adapter("adapter_1")) == 10
    keys = [k for k in tree.state_dict() if "adapter_1" in k]
    assert len(keys) == 5 * 4 and "visual_encoder.0.adapter.adapter_1_down.weight" in keys
