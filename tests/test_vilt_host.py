"""Host logic of the ViLT wrappers on CPU (fp32, no adapters: the adapter arithmetic is CUDA-only):
the dense fast path and the SDPA attention patch reproduce the stock HF forward."""
import logging

import torch

from feddat_b200.modeling.vilt import ViltEncoderWrapper, create_vilt_continual_learner_model
from feddat_b200.synthetic import make_vilt_batch


def _encoder(seed=0):
    from transformers import ViltConfig, ViltModel
    torch.manual_seed(seed)
    return ViltEncoderWrapper(None, ViltModel(ViltConfig()).eval(), torch.device("cpu"))


def test_dense_fast_path_equals_hf_forward():
    enc = _encoder()
    batch = make_vilt_batch(2, text_len=32, image_size=224, seed=3)["encodings"]
    with torch.no_grad():
        ref = enc(**batch)                                           # stock HF path (patch shuffle inside)
        fast = enc(dense_masks=True, **batch)
    assert ref.shape == (2, 768)
    assert (ref - fast).abs().max().item() < 2e-4


def test_sdpa_patch_equals_eager_attention():
    enc = _encoder(1)
    batch = make_vilt_batch(2, text_len=16, image_size=224, seed=4)["encodings"]
    with torch.no_grad():
        a = enc(dense_masks=True, **batch)
        enc._embed_cache = None
        enc.enable_sdpa()
        b = enc(dense_masks=True, **batch)
    assert (a - b).abs().max().item() < 2e-4


def test_learner_keys_and_hooks():
    from feddat_b200.configs.task_configs_fed import task_configs
    model_config = {"encoder_dim": 768, "adapter_config": {"names": ["adapter_0", "adapter_1", "adapter_2"], "device": "cpu"}}
    m = create_vilt_continual_learner_model(logging.getLogger("t"), "random", ["art", "gqa"], model_config,
                                            task_configs, "cpu")
    m.add_adapter()
    keys = list(m.state_dict().keys())
    assert "vilt_encoder.vilt.encoder.layer.11.output.adapter.adapter_1_up.weight" in keys
    assert "vilt_encoder.vilt.encoder.layer.0.output.layer.dense.weight" in keys
    assert "task_layer.gqa.clf_fc1.bias" in keys
    assert len([k for k in keys if "adapter_1" in k]) == 48           # communicated set, SURVEY Appendix B
    m.activate_gating(); m.set_active_adapter("adapter_0")
    ads = m._adapters()
    assert all(a.gating for a in ads)
    assert all(a.adapter_0_down.weight.requires_grad and not a.adapter_1_down.weight.requires_grad for a in ads)
    m.deactivate_gating(); m.set_active_adapter("adapter_1")
    assert all((not a.gating) and a.adapter_1_up.bias.requires_grad and not a.adapter_0_up.bias.requires_grad
               for a in ads)
    assert len(m.get_param_adapter("adapter_1")) == 24


def test_frozen_qkv_backward_equals_autograd():
    """_FrozenQKV (three frozen projections, data gradients accumulated inside the GEMMs) == plain autograd."""
    from feddat_b200.modeling.vilt import _FrozenQKV
    torch.manual_seed(0)
    lin = [torch.nn.Linear(768, 768) for _ in range(3)]
    x1 = torch.randn(2, 5, 768, requires_grad=True)
    x2 = x1.detach().clone().requires_grad_(True)
    gs = [torch.randn(2, 5, 768) for _ in range(3)]
    outs = _FrozenQKV.apply(x1, *[t for m in lin for t in (m.weight.detach(), m.bias.detach())])
    torch.autograd.backward(outs, gs)
    ref = [m(x2) for m in lin]
    torch.autograd.backward(ref, gs)
    for a, b in zip(outs, ref):
        assert torch.allclose(a, b, atol=1e-6)
    assert torch.allclose(x1.grad, x2.grad, atol=1e-5)
