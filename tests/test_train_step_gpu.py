"""Step-level parity: this repo's TaskTrainer.train_step (bf16 backbone, sm_100a DAT / MKD kernels)
against the golden produced by the REFERENCE's own TaskTrainer.train_step + Adapter in fp32
(tests/golden/make_step_golden.py), same seeded init, same batches.  GPU only."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def step_golden():
    from pathlib import Path
    return np.load(Path(__file__).resolve().parent / "golden" / "step_golden.npz")


def build_trainer(model, lr, max_steps, task, temp=3):
    from feddat_b200.modeling.vilt import convert_batch_to_vilt_input_dict
    from feddat_b200.train.accelerator import Accelerator
    from feddat_b200.train.task_trainer import TaskTrainer
    tr = TaskTrainer()
    tr.args = SimpleNamespace(optimizer_mode="dat", encoder_name="vilt", debug=0)
    tr.accelerator = Accelerator(device="cuda")
    tr.device = torch.device("cuda")
    tr.task_key = task
    tr.batch2inputs_converter = convert_batch_to_vilt_input_dict
    tr.loss_criterion = nn.BCEWithLogitsLoss(reduction="mean")
    tr.weight_decay, tr.lr, tr.adam_epsilon = 1e-2, lr, 1e-8
    tr.kl_temp = temp
    tr.max_steps, tr.warmup_ratio = max_steps, 0.1
    return tr


def test_train_step_matches_reference_trainer(step_golden):
    from feddat_b200 import ops
    from feddat_b200.synthetic import make_vilt_batch, to_device
    from feddat_b200.train.prepare import default_args, place_on_gpu, prepare_model
    from feddat_b200.train.task_trainer import get_polynomial_decay_schedule_with_warmup

    seed, rank, steps, max_steps, B, T, H, C = (int(v) for v in step_golden["meta"])
    lr = float(step_golden["lr"])
    torch.manual_seed(seed)
    args = default_args(ordered_cl_tasks=["art"], adapter_rank=rank)
    model = prepare_model(args, place=False)                       # same CPU RNG stream as the golden
    sd0 = {k: v.clone() for k, v in model.state_dict().items() if "adapter" in k or "task_layer" in k}
    place_on_gpu(model)

    tr = build_trainer(model, lr, max_steps, "art")
    # TaskTrainer.train prologue (task_trainer.py:36-45)
    sd = model.state_dict()
    for name in sd:
        if "adapter_1" in name:
            sd[name.replace("adapter_1", "adapter_2")].data.copy_(sd[name].data)
    for n, p in model.named_parameters():
        if "adapter_2" in n:
            p.requires_grad = False
    wrapped = tr.accelerator.prepare(model)
    opt = tr.create_optimizer(wrapped)
    assert sum(len(g["params"]) for g in opt.param_groups) == int(step_golden["n_optimizer_tensors"])
    sched = get_polynomial_decay_schedule_with_warmup(opt, int(max_steps * 0.1), max_steps, lr_end=0, power=1)

    launches0 = ops.launch_count
    wrapped.train()
    report = []
    for step in range(steps):
        batch = to_device(make_vilt_batch(B, T, H, C, seed=seed + step), "cuda")
        loss_0 = tr.train_step(wrapped, step, batch, opt, sched)
        torch.cuda.synchronize()
        got = {n: t.float().cpu().numpy() for n, t in zip(("logits_all", "logits_1", "logits_0"), tr.last_logits)}
        for n in got:
            want = step_golden[f"step{step}/{n}"]
            report.append((f"step{step}/{n}", float(np.linalg.norm(got[n] - want) / np.linalg.norm(want)),
                           float(np.abs(got[n] - want).max() / np.abs(want).max())))
        want_loss = float(step_golden[f"step{step}/loss_0"])
        report.append((f"step{step}/loss_0", abs(loss_0.item() - want_loss) / want_loss, 0.0))
    print("\nstep parity (relative Frobenius, relative max-norm):")
    for name, fro, mx in report:
        print(f"  {name:20s} {fro:.4f} {mx:.4f}")
    assert ops.launch_count > launches0, "the CUDA kernels did not run"
    # Step 0 is pure forward parity (no parameter has moved yet: the first optimizer step of a round
    # has lr = 0 for pass B and the pass-C update only shows from step 1 on): bf16 tolerance 2e-2 --
    # the reference's own arithmetic evaluated in bf16 on the CPU sits 1.2e-2 from its fp32 logits
    # after one forward (measured).
    # From step 1 on the logits also carry AdamW's first updates, which are sign-like
    # (m / (sqrt(v) + eps) = +-1): every gradient entry whose sign differs between the bf16 and the
    # fp32 trajectory moves its weight by lr in the opposite direction, so the trajectories separate
    # by a few percent per step even though the gradients agree to bf16 accuracy (op-level tests).
    # Bars: logits 8e-2 relative Frobenius, task loss 2e-2.
    for name, fro, mx in report:
        if name.endswith("loss_0"):
            assert fro < 2e-2, (name, fro)
        elif name.startswith("step0/"):
            assert fro < 2e-2 and mx < 2e-2, (name, fro, mx)
        else:
            assert fro < 8e-2, (name, fro, mx)

    # parameter movement after 3 steps (AdamW + poly schedule, lr = 0 on the very first optimizer step)
    sd1 = model.state_dict()
    rel = []
    for k in sd0:
        if "adapter_2" in k:
            continue
        want = float(step_golden[f"delta_norm/{k}"])
        got_d = (sd1[k].float().cpu() - sd0[k]).double().norm().item()
        if want > 1e-6:
            rel.append(abs(got_d - want) / want)
        else:
            assert got_d < 1e-5
    print("parameter-movement norms vs reference: median rel err %.4f, max %.4f" % (np.median(rel), max(rel)))
    assert np.median(rel) < 5e-2 and max(rel) < 0.25, (np.median(rel), max(rel))


def test_graphed_train_step_equals_eager():
    """CUDA-graph replay of train_step == the eager train_step (same kernels, same order)."""
    from feddat_b200.synthetic import make_vilt_batch, to_device
    from feddat_b200.train.graphed import GraphedTrainStep
    from feddat_b200.train.prepare import default_args, place_on_gpu, prepare_model
    from feddat_b200.train.task_trainer import get_polynomial_decay_schedule_with_warmup

    def build():
        torch.manual_seed(7)
        model = prepare_model(default_args(ordered_cl_tasks=["art"], adapter_rank=32), place=False)
        place_on_gpu(model)
        tr = build_trainer(model, 1e-3, 20, "art", temp=2.0)
        for n, p in model.named_parameters():
            if "adapter_2" in n:
                p.requires_grad = False
        wrapped = tr.accelerator.prepare(model)
        opt = tr.create_optimizer(wrapped)
        sched = get_polynomial_decay_schedule_with_warmup(opt, 2, 20, lr_end=0, power=1)
        wrapped.train()
        return model, tr, wrapped, opt, sched

    batches = [to_device(make_vilt_batch(2, 16, 224, 100, seed=50 + i), "cuda") for i in range(5)]
    m1, tr1, w1, o1, s1 = build()
    eager = [tr1.train_step(w1, i, batches[i], o1, s1).item() for i in range(5)]
    m2, tr2, w2, o2, s2 = build()
    g = GraphedTrainStep(tr2, w2, o2, s2, batches[0], warmup=2)
    graphed = [g(batches[i]).item() for i in range(5)]
    assert g.graph is not None and g.launches_per_step > 0
    np.testing.assert_allclose(graphed, eager, rtol=2e-3)
    assert s2.last_epoch == s1.last_epoch == 10
    # Parameters: relative Frobenius distance.  Element-wise bars are not meaningful here: the wgrad
    # kernel reduces with fp32 atomics (summation order varies run to run), and Adam turns a gradient
    # that is pure rounding noise into a full +-lr step, so isolated elements legitimately differ by
    # O(lr) between two runs of the SAME code path.
    for (n1, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        if "adapter_0" in n1 or "adapter_1" in n1 or "task_layer" in n1:
            assert ((p1 - p2).norm() / p1.norm().clamp_min(1e-12)).item() < 2e-2, n1


def test_adapter_forward_sees_every_kind_of_parameter_update():
    """Regression: the packed bf16 operands must follow the fp32 masters through updates that do NOT
    bump Tensor._version -- fused AdamW and the reference's ``state_dict()[k].data.copy_`` idiom."""
    from feddat_b200.modeling.adapter import Adapter

    def ref(a, name, x):
        d, u = getattr(a, f"{name}_down"), getattr(a, f"{name}_up")
        bf = lambda t: t.detach().to(torch.bfloat16).float()                     # noqa: E731
        h = torch.relu(x.float() @ bf(d.weight).T + d.bias.detach()).to(torch.bfloat16).float()
        return x.float() + h @ bf(u.weight).T + u.bias.detach()

    torch.manual_seed(3)
    a = Adapter(names=["adapter_0", "adapter_1", "adapter_2"], device="cuda", rank=32)
    for p in a.parameters():
        p.data.normal_(0, 0.05)
    a.set_active_adapter("adapter_1")
    x = torch.randn(300, 768, device="cuda").to(torch.bfloat16)
    params = a._branch_params(("adapter_1",))
    opt = torch.optim.AdamW(params, lr=5e-2, fused=True)

    def check(tag):
        y = a(x, x).float()
        want = ref(a, "adapter_1", x)
        err = ((y - want).abs().max() / want.abs().max()).item()
        assert err < 1e-2, (tag, err)
        return y

    y0 = check("initial")
    a(x, x).float().square().mean().backward()
    opt.step()                                                                   # no _version bump
    y1 = check("after fused AdamW step")
    assert (y1 - y0).abs().max().item() > 1e-2                                  # the update is visible
    sd = a.state_dict()
    sd["adapter_1_up.weight"].data.copy_(torch.randn_like(sd["adapter_1_up.weight"]) * 0.05)
    y2 = check("after state_dict().data.copy_")
    assert (y2 - y1).abs().max().item() > 1e-2


def test_fused_layernorm_layer_equals_stock_hf_layer():
    """fast_vilt_layer_forward (fused add + LayerNorm kernels) == the stock HF ViltLayer.forward on the
    same bf16 layer: output and input gradient."""
    from feddat_b200.train.prepare import default_args, place_on_gpu, prepare_model
    torch.manual_seed(3)
    model = prepare_model(default_args(ordered_cl_tasks=["art"], adapter_rank=32), place=False)
    place_on_gpu(model)
    model.activate_gating(); model.set_active_adapter("adapter_0")
    layer = model.vilt_encoder.vilt.encoder.layer[5]
    h0 = torch.randn(4, 37, 768, device="cuda").to(torch.bfloat16)
    outs = []
    for fused in (True, False):
        h = h0.clone().requires_grad_(True)
        y = layer(h, None)[0] if fused else type(layer).forward(layer, h, None)[0]
        y.float().square().mean().backward()
        outs.append((y.detach().float(), h.grad.float()))
    (y1, g1), (y2, g2) = outs
    assert ((y1 - y2).abs().max() / y2.abs().max()).item() < 1e-2
    assert ((g1 - g2).abs().max() / g2.abs().max()).item() < 2e-2


def test_graphed_step_with_prefetch_equals_direct_batches():
    """GraphedTrainStep.prefetch() (next batch staged host->device on a copy stream, consumed by the next
    call without a batch) == feeding the same batches directly."""
    from feddat_b200.synthetic import make_vilt_batch
    from feddat_b200.train.graphed import GraphedTrainStep
    from feddat_b200.train.prepare import default_args, place_on_gpu, prepare_model
    from feddat_b200.train.task_trainer import get_polynomial_decay_schedule_with_warmup
    from feddat_b200.synthetic import to_device

    def build():
        torch.manual_seed(11)
        model = prepare_model(default_args(ordered_cl_tasks=["art"], adapter_rank=32), place=False)
        place_on_gpu(model)
        tr = build_trainer(model, 1e-3, 20, "art", temp=2.0)
        for n, p in model.named_parameters():
            if "adapter_2" in n:
                p.requires_grad = False
        wrapped = tr.accelerator.prepare(model)
        opt = tr.create_optimizer(wrapped)
        sched = get_polynomial_decay_schedule_with_warmup(opt, 2, 20, lr_end=0, power=1)
        wrapped.train()
        return tr, wrapped, opt, sched

    host = [make_vilt_batch(2, 16, 224, 100, seed=70 + i, pin=True) for i in range(5)]
    example = to_device(host[0], "cuda")
    tr1, w1, o1, s1 = build()
    g1 = GraphedTrainStep(tr1, w1, o1, s1, example, warmup=1)
    direct = [g1(host[i]).item() for i in range(5)]
    tr2, w2, o2, s2 = build()
    g2 = GraphedTrainStep(tr2, w2, o2, s2, example, warmup=1)
    staged = []
    g2.prefetch(host[0])
    for i in range(5):
        loss = g2()
        if i + 1 < 5:
            g2.prefetch(host[i + 1])
        staged.append(loss.item())
    np.testing.assert_allclose(staged, direct, rtol=2e-3)
