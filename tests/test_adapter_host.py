"""Host-side contract of the drop-in Adapter (reference adapter.py): names, shapes, requires_grad
toggling, error behaviour.  CPU only -- the arithmetic itself is GPU-tested."""
import numpy as np
import pytest
import torch

from feddat_b200._lib import FeddatError
from feddat_b200.modeling.adapter import Adapter

NAMES = ["adapter_0", "adapter_1", "adapter_2"]


def test_state_dict_keys_and_shapes_match_reference():
    ad = Adapter(names=NAMES, device="cpu", model_dim=768)           # reference default: r = 768 // 16
    sd = ad.state_dict()
    want = {}
    for n in NAMES:
        want[f"{n}_down.weight"] = (48, 768)
        want[f"{n}_down.bias"] = (48,)
        want[f"{n}_up.weight"] = (768, 48)
        want[f"{n}_up.bias"] = (768,)
    assert {k: tuple(v.shape) for k, v in sd.items()} == want
    assert all(float(sd[f"{n}_down.bias"].abs().max()) == 0.0 for n in NAMES)      # adapter.py:13-14
    assert 0.015 < float(sd["adapter_0_down.weight"].std()) < 0.025                # N(0, 0.02)


@pytest.mark.parametrize("case,mode", [("single_r16", "single"), ("gating_r16", "gating")])
def test_requires_grad_toggling_matches_reference(golden, case, mode):
    ad = Adapter(names=NAMES, device="cpu", model_dim=768, adapter_reduction_factor=48)
    if mode == "single":
        ad.deactivate_gating(); ad.set_active_adapter("adapter_1")
    else:
        ad.activate_gating(); ad.set_active_adapter("adapter_0")
    got = np.array([int(p.requires_grad) for _, p in sorted(ad.named_parameters())])
    assert np.array_equal(got, golden[f"adapter/{case}/requires_grad"])


def test_rank_plumbing_and_validation():
    assert Adapter(NAMES, "cpu", adapter_reduction_factor=6).rank == 128
    assert Adapter(NAMES, "cpu", rank=512).adapter_0_up.weight.shape == (768, 512)
    with pytest.raises(FeddatError):
        Adapter(NAMES, "cpu", rank=24)                 # not a multiple of 16: no fallback path
    with pytest.raises(FeddatError):
        Adapter(NAMES, "cpu", model_dim=1024)
    with pytest.raises(ValueError):
        Adapter(NAMES, "cpu", activation="tanh")


def test_cpu_forward_fails_loudly():
    ad = Adapter(NAMES, "cpu")
    ad.set_active_adapter("adapter_1")
    x = torch.zeros(2, 3, 768)
    with pytest.raises(FeddatError, match="no fallback"):
        ad(x, x)


def test_forward_without_active_adapter_raises_like_reference():
    ad = Adapter(NAMES, "cpu")
    with pytest.raises((AttributeError, FeddatError)):
        ad(torch.zeros(1, 1, 768), torch.zeros(1, 1, 768))

