"""Host-side checks of the ALBEF path (no GPU): state-dict key compatibility with the reference's model
(the key list in the golden was read off the executed reference modules), parameter selection of
prepare_model (main.py:127-163,248-250), the synthetic batch contract, and the oracle of the fused MKD head
against the torch ops the reference calls."""
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

import oracle
from feddat_b200.synthetic import make_albef_batch
from feddat_b200.train.prepare import default_args, prepare_model
from tests.golden_inputs import ALBEF_GOLDEN_CFG

GOLD = Path(__file__).resolve().parent / "golden" / "albef_step_golden.npz"


def small_albef(rank=64):
    args = default_args(encoder_name="albef_no_distill", ordered_cl_tasks=["art"], adapter_rank=rank,
                        image_size=ALBEF_GOLDEN_CFG["image_res"], vit_depth=ALBEF_GOLDEN_CFG["vit_depth"],
                        decoder_layers=ALBEF_GOLDEN_CFG["decoder_layers"], bert_overrides=ALBEF_GOLDEN_CFG["bert_config"])
    return prepare_model(args, place=False), args


def test_state_dict_keys_equal_the_reference_model():
    gold = np.load(GOLD)
    model, args = small_albef()
    assert sorted(model.state_dict().keys()) == list(gold["r64/state_dict_keys"])
    # communicated = adapter_1 of all 2 + 2 + 1 sites; personal = LM head + adapter_0 / adapter_2
    assert len(model.comm_state_dict_names) == 5 * 4 and all("adapter_1" in n for n in model.comm_state_dict_names)
    assert args.personal_params_names == [".cls.", "adapter_0", "adapter_2"]
    trainable = [n for n, p in model.named_parameters() if p.requires_grad]
    assert all(("adapter" in n) or (".cls." in n) for n in trainable)
    assert sum(".cls." in n for n in trainable) == 6


def test_full_size_architecture_matches_the_reference_counts():
    """ALBEF at full depth: 12 ViT + 12 text-encoder + 6 text-decoder sites."""
    args = default_args(encoder_name="albef_no_distill", ordered_cl_tasks=["art"], adapter_rank=256, image_size=384)
    model = prepare_model(args, place=False)
    ads = model._adapters()
    assert len(ads) == 30
    assert model.albef_model.albef.visual_encoder.pos_embed.shape == (1, 577, 768)
    layers = model.albef_model.albef.text_encoder.encoder.layer
    assert [l.has_cross_attention for l in layers] == [False] * 6 + [True] * 6
    assert all(l.has_cross_attention for l in model.albef_model.albef.text_decoder.bert.encoder.layer)
    n_comm = sum(p.numel() for n, p in model.named_parameters() if "adapter_1" in n)
    assert n_comm == 30 * (2 * 768 * 256 + 256 + 768)            # 11.8 M floats = 47 MB per round


def test_synthetic_albef_batch_contract():
    b = make_albef_batch(6, image_size=64, seed=3)
    n_seq = sum(b["n"])
    assert b["answer_ids"].shape == (n_seq, 6) and b["weights"].shape == (n_seq,)
    assert torch.allclose(b["weights"].sum(), torch.tensor(6.0))        # weights 1 / k_b sum to the batch size
    assert (b["answer_ids"][:, 0] == 101).all() and (b["question_ids"][:, 0] == 101).all()
    last = b["answer_mask"].sum(1) - 1
    assert (b["answer_ids"].gather(1, last[:, None]) == 102).all()      # every answer ends in [SEP]
    assert sum(make_albef_batch(6, 64, seed=4)["n"]) == n_seq                      # static shapes across batches
    assert sum(make_albef_batch(16, 64, seed=7)["n"]) == sum(make_albef_batch(16, 64, seed=12)["n"])
    assert b["answer_index"].tolist() == [i for i, k in enumerate(b["n"]) for _ in range(k)]


def test_mkd_ce_oracle_equals_the_reference_torch_expression():
    """oracle.mkd_ce_total == (answer loss of xbert.py:1287-1297 * weights / B + kl_loss(logits[:, :-1], teacher)) / 2
    evaluated with the torch ops the reference calls, values and d/dscores."""
    rng = np.random.default_rng(0)
    n, La, C, B, T = 5, 4, 3201, 3, 2.0
    s = (rng.standard_normal((n, La, C)) * 2).astype(np.float32)
    t = (rng.standard_normal((n, La - 1, C)) * 2).astype(np.float32)
    lab = rng.integers(0, C, (n, La))
    lab[0, 2:] = -100
    lab[3, 3] = -100
    w = rng.random(n).astype(np.float32)
    st = torch.tensor(s, requires_grad=True)
    shifted = st[:, :-1, :].contiguous()
    labs = torch.tensor(lab)[:, 1:].contiguous()
    lm = torch.nn.CrossEntropyLoss(reduction="none")(shifted.view(-1, C), labs.view(-1)).view(n, -1).sum(1)
    task = (torch.tensor(w) * lm).sum() / B
    logits = st[:, :-1, :].contiguous()
    kl = F.kl_div(F.log_softmax(logits / T, dim=-1), F.softmax(torch.tensor(t) / T, dim=-1), reduction="batchmean") * T ** 2
    L = (task + kl) / 2
    L.backward()
    Lo, klo, tasko, g = oracle.mkd_ce_total(s, t, lab, w, B, T)
    assert abs(Lo - L.item()) < 1e-4 and abs(klo - kl.item()) < 1e-4 and abs(tasko - task.item()) < 1e-4
    assert np.abs(g - st.grad.numpy()).max() < 1e-6
