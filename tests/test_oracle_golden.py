"""Pins the numpy oracle against the golden vectors produced by executing the reference's own code
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

import oracle
from tests.golden_inputs import adapter_inputs, bert_inputs, branch, fedavg_inputs, kl_inputs

ADAPTER_CASES = ["single_r16", "gating_r16", "single_r48", "gating_r48", "gating_r128"]


def _close(a, b, rtol=2e-5, atol=2e-5):
    np.testing.assert_allclose(np.asarray(a, np.float64), np.asarray(b, np.float64), rtol=rtol, atol=atol)


@pytest.mark.parametrize("case", ADAPTER_CASES)
def test_adapter_forward_backward_matches_reference(golden, case):
    w, x, g, r, gating = adapter_inputs(golden[f"adapter/{case}/meta"])
    # reference modes: single = adapter_1 alone (task_trainer.py:290-291),
    # gating = adapter_0 (trainable) + adapter_2 (frozen) (task_trainer.py:311-312, adapter.py:133-146)
    branches = [branch(w, "adapter_0"), branch(w, "adapter_2")] if gating else [branch(w, "adapter_1")]
    y = oracle.adapter_forward(x, x, branches, gating)
    _close(y, golden[f"adapter/{case}/y"])
    dx, grads = oracle.adapter_backward(x, g, branches, gating, residual_is_input=True)
    _close(dx, golden[f"adapter/{case}/dx"], rtol=1e-4, atol=1e-4)
    d_down_w, d_down_b, d_up_w, d_up_b = grads[0]     # the trainable branch is first in both modes
    _close(d_down_b, golden[f"adapter/{case}/d_down_b"], rtol=1e-4, atol=1e-4)
    _close(d_up_b, golden[f"adapter/{case}/d_up_b"], rtol=1e-4, atol=1e-4)
    if f"adapter/{case}/d_down_w" in golden:
        _close(d_down_w, golden[f"adapter/{case}/d_down_w"], rtol=1e-4, atol=1e-4)
        _close(d_up_w, golden[f"adapter/{case}/d_up_w"], rtol=1e-4, atol=1e-4)


def test_bert_site_wrapper_matches_reference(golden):
    w, ffn, x, lnw, lnb, r = bert_inputs(golden["bert/gating_r16/meta"])
    y = oracle.adapter_layer_forward_bert(ffn, x, lnw, lnb, 1e-12,
                                          [branch(w, "adapter_0"), branch(w, "adapter_2")], gating=True)
    _close(y, golden["bert/gating_r16/y"], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("case", ["vilt_T3", "vilt_T2", "wide_T3"])
def test_kl_loss_matches_reference(golden, case):
    a, b, temp = kl_inputs(golden[f"kl/{case}/meta"])
    loss, grad = oracle.kl_loss(a, b, temp, with_grad=True)
    _close(loss, golden[f"kl/{case}/loss"], rtol=1e-5, atol=1e-6)
    _close(grad, golden[f"kl/{case}/grad"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("case", ["vilt_T3", "vilt_T2"])
def test_mkd_total_matches_reference(golden, case):
    a, b, temp = kl_inputs(golden[f"kl/{case}/meta"])
    total, kl, task, grad = oracle.mkd_total(a, b, golden[f"mkd/{case}/target"], temp)
    _close(task, golden[f"mkd/{case}/task"], rtol=1e-5, atol=1e-5)
    _close(total, golden[f"mkd/{case}/total"], rtol=1e-5, atol=1e-5)
    _close(grad, golden[f"mkd/{case}/grad"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("case", ["equal3", "weighted3", "equal8"])
def test_fedavg_bit_exact(golden, case):
    keys, clients, nums = fedavg_inputs(golden[f"fedavg/{case}/meta"])
    for k in keys:
        out = oracle.get_average_net([c[k] for c in clients], nums)
        assert out.dtype == np.float32
        assert np.array_equal(out, golden[f"fedavg/{case}/{k}"]), "FedAvg must be bit-exact in fp32"
