"""Golden vectors of the ALBEF adapter-injection sites, produced by EXECUTING THE REFERENCE'S OWN
``Block`` (src/modeling/models/vit.py:78-110) and ``BertOutput`` (src/modeling/models/xbert.py:428-445)
in fp32 on CPU (run in the build container, where /root/reference is mounted).

    python tests/golden/make_albef_site_golden.py

Load-time shims only (nothing is modified or copied): a stub ``timm``
(PatchEmbed / trunc_normal_ / DropPath are not used by Block), two helpers re-exported into
``transformers.modeling_utils`` for xbert's imports, and the ``.to('cuda')`` of adapter.py:144,160
redirected to the input's device.  Weights are NOT stored: ``tests/golden_inputs.fill_params`` fills any
module deterministically by parameter name, here and in the GPU test.
"""
from __future__ import annotations

import contextlib
import importlib
import importlib.machinery
import io
import sys
import types
from functools import partial
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent
sys.path.insert(0, str(OUT.parent.parent))
from tests.golden_inputs import albef_site_gout, albef_site_inputs, fill_params  # noqa: E402

NAMES = ["adapter_0", "adapter_1", "adapter_2"]


def load_reference_modules():
    for name in ("timm", "timm.models", "timm.models.vision_transformer", "timm.models.registry",
                 "timm.models.layers"):
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)      # transformers probes find_spec("timm")
        m.__path__ = []
        sys.modules.setdefault(name, m)
    vt = sys.modules["timm.models.vision_transformer"]
    vt._cfg = lambda **kw: {}
    vt.PatchEmbed = nn.Identity
    sys.modules["timm.models.registry"].register_model = lambda f: f
    sys.modules["timm.models.layers"].trunc_normal_ = nn.init.trunc_normal_
    sys.modules["timm.models.layers"].DropPath = nn.Identity
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu
    for n in ("apply_chunking_to_forward", "prune_linear_layer"):
        if not hasattr(mu, n):
            setattr(mu, n, getattr(pu, n))
    if not hasattr(mu, "find_pruneable_heads_and_indices"):
        mu.find_pruneable_heads_and_indices = None
    pkg = types.ModuleType("refmodels")
    pkg.__path__ = [str(REF / "src/modeling/models")]
    sys.modules["refmodels"] = pkg
    src = (REF / "src/modeling/models/adapter.py").read_text()
    assert src.count(".to('cuda')") == 2
    ad = types.ModuleType("refmodels.adapter")
    exec(compile(src.replace(".to('cuda')", ".to(hidden_states.device)"), "adapter.py", "exec"), ad.__dict__)
    sys.modules["refmodels.adapter"] = ad
    return importlib.import_module("refmodels.vit"), importlib.import_module("refmodels.xbert")


def run(module, call, params_filter):
    """forward + backward; returns outputs, input grads and the adapter-parameter grads."""
    module.zero_grad()
    out, inputs = call()
    out.backward(torch.from_numpy(albef_site_gout(out.shape)))
    res = {"out": out.detach().numpy()}
    for k, t in inputs.items():
        if k != "h":                       # d(ffn input) is plain frozen-dense backprop: not stored
            res[f"d_{k}"] = t.grad.numpy().copy()
    for n, p in module.named_parameters():
        if params_filter(n) and p.grad is not None:
            res[f"grad/{n}"] = p.grad.numpy().copy()
    return res


def main():
    vit, xbert = load_reference_modules()
    rank = 64
    acfg = {"names": NAMES, "device": "cpu", "adapter_reduction_factor": 768 // rank}
    out = {}
    with contextlib.redirect_stdout(io.StringIO()):
        block = vit.Block(dim=768, num_heads=12, mlp_ratio=4, qkv_bias=True,
                          norm_layer=partial(nn.LayerNorm, eps=1e-6), adapter_config=dict(acfg))
        cfg = types.SimpleNamespace(intermediate_size=3072, hidden_size=768, layer_norm_eps=1e-12,
                                    hidden_dropout_prob=0.0, adapter_config=dict(acfg))
        bout = xbert.BertOutput(cfg)
    fill_params(block, seed=11)
    fill_params(bout, seed=12)
    inp = albef_site_inputs()
    for mode in ("single_adapter_1", "gating"):
        for mod in (block.adapter, bout.adapter):
            if mode == "gating":
                mod.activate_gating(); mod.set_active_adapter("adapter_0")
            else:
                mod.deactivate_gating(); mod.set_active_adapter("adapter_1")

        def call_block():
            x = torch.from_numpy(inp["vit_x"]).requires_grad_(True)
            return block(x), {"x": x}

        def call_bout():
            h = torch.from_numpy(inp["bert_h"]).requires_grad_(True)
            x = torch.from_numpy(inp["bert_x"]).requires_grad_(True)
            return bout(h, x), {"h": h, "x": x}

        for name, mod, call in (("vit_block", block, call_block), ("bert_output", bout, call_bout)):
            for k, v in run(mod, call, lambda n: "adapter" in n).items():
                out[f"{name}/{mode}/{k}"] = v.astype(np.float32)
    out["meta_rank"] = np.array([rank])
    np.savez_compressed(OUT / "albef_site_golden.npz", **out)
    print("wrote", OUT / "albef_site_golden.npz", {k: v.shape for k, v in list(out.items())[:6]}, len(out), "arrays")


if __name__ == "__main__":
    main()
