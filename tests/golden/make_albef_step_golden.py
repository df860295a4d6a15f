"""Step-level golden of the ALBEF path (BASELINE configs[2]), produced by EXECUTING THE REFERENCE'S OWN CODE on
CPU in fp32: ``TaskTrainer.train_step`` / ``create_optimizer`` / ``kl_loss`` (src/train/visionlanguage_tasks/
task_trainer.py, imported unmodified) driving ``ALBEF.forward`` (src/modeling/models/albef_model.py:69-145, the
function taken from the source by AST and bound to a holder: ``ALBEF.__init__`` hard-codes a private BERT path)
over the reference's vendored ``VisionTransformer`` / ``BertModel`` / ``BertLMHeadModel`` (vit.py, xbert.py) with
the reference ``Adapter`` at all sites.  Load-time shims: SURVEY.md Appendix A.2 (tests/golden/
make_albef_site_golden.py::load_reference_modules), plus a real Conv2d PatchEmbed for the stubbed timm one.

Reduced depth so the fixture stays small and the CPU run takes seconds (the arithmetic per layer is what is
pinned; every layer type is present): image 64 x 64 (17 ViT tokens), ViT depth 2, question encoder 2 layers with
fusion_layer 1 (one text-only layer, one cross-attention layer), answer decoder 1 layer, vocabulary 3 200 (> 3 000:
kl_loss's last-dim branch, task_trainer.py:507-509), dropout 0 (the CPU and GPU random streams cannot match).
Ranks 64 and 256 (R = 512 in gating mode: two segment launches per site on the GPU side).  Weights are NOT
stored: ``tests/golden_inputs.fill_params`` fills both sides by parameter name.

    python tests/golden/make_albef_step_golden.py      # writes tests/golden/albef_step_golden.npz
"""
from __future__ import annotations

import ast
import sys
import types
from argparse import Namespace
from functools import partial
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[2]
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))

from make_albef_site_golden import NAMES, load_reference_modules  # noqa: E402
from tests.golden_inputs import ALBEF_GOLDEN_CFG, albef_golden_batch, fill_params, grad_sketch  # noqa: E402

STEPS, MAX_STEPS, LR, TEMP = 2, 10, 1e-3, 2.0


class PatchEmbed(nn.Module):
    """What vit.py:144-146 uses of timm's PatchEmbed."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        self.num_patches = (img_size // patch_size) ** 2
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


def reference_albef_forward():
    """ALBEF.forward (albef_model.py:69-156) as a plain function."""
    path = REF / "src/modeling/models/albef_model.py"
    tree = ast.parse(path.read_text())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "ALBEF")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "forward")
    mod = types.ModuleType("ref_albef_forward")
    mod.__dict__.update(torch=torch, F=F, nn=nn, np=np)
    exec(compile(ast.Module(body=[fn], type_ignores=[]), str(path), "exec"), mod.__dict__)
    return mod.forward


def build_reference(rank):
    vit, xbert = load_reference_modules()
    vit.PatchEmbed = PatchEmbed
    xbert.BertPreTrainedModel.init_weights = lambda self: self.apply(self._init_weights)
    xbert.BertModel.get_head_mask = lambda self, hm, n, *a, **k: [None] * n
    c = ALBEF_GOLDEN_CFG
    adapter_config = {"names": NAMES, "device": "cpu", "adapter_reduction_factor": 768 // rank}
    from transformers import BertConfig
    enc = BertConfig(**c["bert_config"])
    dec = BertConfig(**c["bert_config"])
    dec.fusion_layer = 0                                             # albef_model.py:32-33
    dec.num_hidden_layers = c["decoder_layers"]
    enc.adapter_config = adapter_config
    dec.adapter_config = adapter_config

    class Holder(nn.Module):                                         # stands in for ALBEF.__init__ (:13-57)
        def __init__(self):
            super().__init__()
            import contextlib
            import io
            with contextlib.redirect_stdout(io.StringIO()):
                self.visual_encoder = vit.VisionTransformer(
                    img_size=c["image_res"], patch_size=16, embed_dim=768, depth=c["vit_depth"], num_heads=12,
                    mlp_ratio=4, qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6), adapter_config=adapter_config)
                self.text_encoder = xbert.BertModel(enc, add_pooling_layer=False)
                self.text_decoder = xbert.BertLMHeadModel(dec)
            self.tokenizer = Namespace(pad_token_id=0)
            self.distill = False

    Holder.forward = reference_albef_forward()

    class Wrapper(nn.Module):                                        # albef.py:24-72 with pre-tokenised inputs
        def __init__(self):
            super().__init__()
            self.albef = Holder()

        def forward(self, batch):
            q = Namespace(input_ids=batch["question_ids"], attention_mask=batch["question_mask"])
            a = Namespace(input_ids=batch["answer_ids"], attention_mask=batch["answer_mask"])
            loss, logits = self.albef(image=batch["images"], question=q, answer=a, train=True, alpha=batch["alpha"],
                                      k=batch["n"], weights=batch["weights"])
            return [loss, logits]

    class Learner(nn.Module):                                        # albef.py:104-183
        def __init__(self):
            super().__init__()
            self.albef_model = Wrapper()

        def _ads(self):
            return [m for m in self.modules() if type(m).__name__ == "Adapter"]

        def set_active_adapter(self, name):
            [a.set_active_adapter(name) for a in self._ads()]

        def activate_gating(self):
            [a.activate_gating() for a in self._ads()]

        def deactivate_gating(self):
            [a.deactivate_gating() for a in self._ads()]

        def forward(self, task_key, batch):
            return self.albef_model(batch)

    return Learner()


def prepare(learner):
    """main.py:138-139,157-159,248-250 + task_trainer.py:36-45."""
    for p in learner.parameters():
        p.requires_grad = False
    for n, p in learner.named_parameters():
        if "adapter" in n or ".cls." in n:
            p.requires_grad = True
    sd = learner.state_dict()
    for name in sd:
        if "adapter_1" in name:
            sd[name.replace("adapter_1", "adapter_2")].data.copy_(sd[name].data)
    for n, p in learner.named_parameters():
        if "adapter_2" in n:
            p.requires_grad = False


def main():
    sys.path.insert(0, str(REF))
    import src.train.visionlanguage_tasks.task_trainer as ref_tt
    from transformers import get_polynomial_decay_schedule_with_warmup
    gold = {"meta": np.array([STEPS, MAX_STEPS]), "lr": np.array(LR), "temp": np.array(TEMP)}
    for rank in (64, 256):
        learner = build_reference(rank)
        fill_params(learner, seed=21)
        prepare(learner)
        learner.train()
        n_keys = len(learner.state_dict())
        gold[f"r{rank}/state_dict_keys"] = np.array(sorted(learner.state_dict().keys()))
        # ---- plain forwards (loss + logits) in both modes
        batch = albef_golden_batch(0)
        for mode in ("gating", "adapter_1"):
            if mode == "gating":
                learner.activate_gating(); learner.set_active_adapter("adapter_0")
            else:
                learner.deactivate_gating(); learner.set_active_adapter("adapter_1")
            with torch.no_grad():
                loss, logits = learner("art", dict(batch, train=True))
            gold[f"r{rank}/fwd/{mode}/loss"] = np.array(loss.item())
            gold[f"r{rank}/fwd/{mode}/logits"] = logits.numpy().astype(np.float32)

        # ---- the reference trainer's train_step (the mode switches above toggled requires_grad: start again
        # from the flags prepare_model leaves, main.py:157-159)
        prepare(learner)

        class Wrap(nn.Module):
            def __init__(self, m):
                super().__init__()
                self.module = m

            def forward(self, *a, **k):
                return self.module(*a, **k)

        grads = []

        def backward(loss):
            loss.backward()
            grads.append({n: p.grad.detach().clone() for n, p in learner.named_parameters() if p.grad is not None})

        tr = ref_tt.TaskTrainer()
        tr.args = Namespace(optimizer_mode="dat", encoder_name="albef_no_distill")
        tr.accelerator = Namespace(device=torch.device("cpu"), backward=backward)
        tr.device, tr.task_key = torch.device("cpu"), "art"
        tr.batch2inputs_converter = lambda b: dict(b)
        tr.weight_decay, tr.lr, tr.adam_epsilon = 1e-2, LR, 1e-8
        tr.kl_criterion = lambda out, tgt: ref_tt.kl_loss(out, tgt, temp=TEMP)
        wrapped = Wrap(learner)
        opt = tr.create_optimizer(wrapped)
        sched = get_polynomial_decay_schedule_with_warmup(opt, num_warmup_steps=int(MAX_STEPS * 0.1),
                                                          num_training_steps=MAX_STEPS, lr_end=0, power=1)
        gold[f"r{rank}/n_optimizer_tensors"] = np.array(sum(len(g["params"]) for g in opt.param_groups))
        for step in range(STEPS):
            batch = albef_golden_batch(step)
            seen = []
            h = learner.albef_model.register_forward_hook(lambda m, i, o: seen.append((o[0].detach().clone(), o[1].detach().clone())))
            grads.clear()
            loss_0 = tr.train_step(wrapped, step, batch, opt, sched)
            h.remove()
            gold[f"r{rank}/step{step}/loss_0"] = np.array(loss_0.item())
            for name, (ls, lg) in zip(("all", "1", "0"), seen):
                gold[f"r{rank}/step{step}/task_loss_{name}"] = np.array(ls.item())
                if step == 0:          # later steps: losses and gradient norms only (fixture size)
                    gold[f"r{rank}/step{step}/logits_{name}"] = lg.numpy().astype(np.float32)
            for tag, gd in zip(("B", "C"), grads):
                for n, g in gd.items():
                    gold[f"r{rank}/step{step}/grad{tag}/norm/{n}"] = np.array(g.double().norm().item())
                    if step == 0:
                        gold[f"r{rank}/step{step}/grad{tag}/sketch/{n}"] = grad_sketch(n, g.float().numpy())
            print(f"[r={rank}] step {step}: loss_0 = {loss_0.item():.6f}; {n_keys} state-dict keys; "
                  f"{len(grads[0])} / {len(grads[1])} gradient tensors after pass B / C")
    np.savez_compressed(ROOT / "tests" / "golden" / "albef_step_golden.npz", **gold)
    print("wrote albef_step_golden.npz with", len(gold), "arrays")


if __name__ == "__main__":
    main()
