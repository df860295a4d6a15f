"""Regenerates the seeded inputs of tests/golden/op_golden.npz (same PCG64 streams as
tests/golden/make_golden.py, so the fixture only has to carry the reference's OUTPUTS)."""
import numpy as np

NAMES = ("adapter_0", "adapter_1", "adapter_2")


def rng_weights(rng, r, d=768):
    w = {}
    for n in NAMES:
        w[f"{n}_down.weight"] = (rng.standard_normal((r, d)) * 0.05).astype(np.float32)
        w[f"{n}_down.bias"] = (rng.standard_normal((r,)) * 0.1).astype(np.float32)
        w[f"{n}_up.weight"] = (rng.standard_normal((d, r)) * 0.05).astype(np.float32)
        w[f"{n}_up.bias"] = (rng.standard_normal((d,)) * 0.1).astype(np.float32)
    return w


def branch(w, name):
    return (w[f"{name}_down.weight"], w[f"{name}_down.bias"], w[f"{name}_up.weight"], w[f"{name}_up.bias"])


def adapter_inputs(meta, d=768):
    seed, r, b, s, gating = (int(v) for v in meta)
    rng = np.random.default_rng(seed)
    w = rng_weights(rng, r, d)
    x = rng.standard_normal((b, s, d)).astype(np.float32)
    g = rng.standard_normal((b, s, d)).astype(np.float32)
    return w, x, g, r, bool(gating)


def bert_inputs(meta, d=768):
    seed, r, b, s, gating = (int(v) for v in meta)
    rng = np.random.default_rng(seed)
    w = rng_weights(rng, r, d)
    ffn = rng.standard_normal((b, s, d)).astype(np.float32)
    x = rng.standard_normal((b, s, d)).astype(np.float32)
    lnw = (1.0 + 0.1 * rng.standard_normal(d)).astype(np.float32)
    lnb = (0.1 * rng.standard_normal(d)).astype(np.float32)
    return w, ffn, x, lnw, lnb, r


def kl_inputs(meta):
    seed, temp, shape = int(meta[0]), float(meta[1]), tuple(int(v) for v in meta[2:])
    rng = np.random.default_rng(seed)
    a = (rng.standard_normal(shape) * 2).astype(np.float32)
    b = (rng.standard_normal(shape) * 2).astype(np.float32)
    return a, b, temp


def fedavg_inputs(meta):
    seed, nums = int(meta[0]), [int(v) for v in meta[1:]]
    rng = np.random.default_rng(seed)
    keys = ["a.adapter_1_down.weight", "a.adapter_1_up.bias"]
    clients = [{k: rng.standard_normal(257 if "bias" in k else (16, 33)).astype(np.float32) for k in keys}
               for _ in nums]
    return keys, clients, nums


def fill_params(module, seed):
    """Deterministic fill of every parameter of ``module`` BY NAME (sorted), independent of the
    construction order and of torch's RNG: adapter weights ~ N(0, .05), backbone weights ~ N(0, .02) (BERT
    init; larger values make the frozen block amplify the operator's bf16 rounding in d_x), LayerNorm
    weights 1 + N(0, .1), biases ~ N(0, .1) (backbone: .02).  Used identically by tests/golden/make_albef_site_golden.py (on the reference modules)
    and by the GPU test (on this repo's modules), so the fixture need not carry 7 M weights."""
    import zlib

    import torch
    for name, p in sorted(module.named_parameters(), key=lambda kv: kv[0]):
        rng = np.random.default_rng([seed, zlib.crc32(name.encode())])
        ad = "adapter" in name
        if p.dim() == 1:
            if "norm" in name.lower() and name.endswith("weight"):
                v = 1.0 + rng.standard_normal(p.shape) * 0.1
            else:
                v = rng.standard_normal(p.shape) * (0.1 if ad else 0.02)
        else:
            v = rng.standard_normal(p.shape) * (0.05 if ad else 0.02)
        with torch.no_grad():
            p.copy_(torch.from_numpy(v.astype(np.float32)))


def albef_site_inputs():
    rng = np.random.default_rng(2024)
    return {"vit_x": rng.standard_normal((2, 21, 768)).astype(np.float32),       # 42 rows: one ragged tile
            "bert_h": rng.standard_normal((2, 13, 3072)).astype(np.float32),
            "bert_x": rng.standard_normal((2, 13, 768)).astype(np.float32)}


def albef_site_gout(shape):
    return np.random.default_rng(99).standard_normal(tuple(shape)).astype(np.float32)


def grad_sketch(name, g, cols=8):
    """Compact fingerprint of a gradient tensor for the step-level goldens: 1-D tensors in full, 2-D tensors
    as G @ Omega with a seeded N(0, 1) matrix Omega [in_features, cols] (keyed by the parameter name, so both
    sides of a comparison draw the same Omega).  A random projection preserves relative Frobenius distances
    in expectation, so comparing sketches bounds the distance of the full gradients."""
    import zlib
    g = np.asarray(g, np.float32)
    if g.ndim < 2:
        return g.copy()
    g2 = g.reshape(g.shape[0], -1)
    rng = np.random.default_rng([77, zlib.crc32(name.encode())])
    omega = rng.standard_normal((g2.shape[1], cols)).astype(np.float32)
    return g2 @ omega


# ------------------------------------------------------------------------------------------------ ALBEF step golden
ALBEF_GOLDEN_CFG = {
    "image_res": 64, "vit_depth": 2, "decoder_layers": 1, "distill": False,
    "bert_config": {"attention_probs_dropout_prob": 0.0, "hidden_act": "gelu", "hidden_dropout_prob": 0.0,
                    "hidden_size": 768, "initializer_range": 0.02, "intermediate_size": 3072, "layer_norm_eps": 1e-12,
                    "max_position_embeddings": 512, "num_attention_heads": 12, "num_hidden_layers": 2,
                    "pad_token_id": 0, "type_vocab_size": 2, "vocab_size": 3200, "fusion_layer": 1,
                    "encoder_width": 768},
}


def albef_golden_batch(step):
    """B = 2 images, questions of 8 / 6 tokens (padded to 8), k = [1, 2] answers of 5 / 4 / 3 tokens (padded to 5)."""
    import torch
    rng = np.random.default_rng(300 + step)
    images = torch.from_numpy(rng.standard_normal((2, 3, 64, 64)).astype(np.float32))
    q = np.zeros((2, 8), np.int64)
    qm = np.zeros((2, 8), np.int64)
    for b, n in enumerate((8, 6)):
        q[b, 0], q[b, n - 1] = 101, 102
        q[b, 1:n - 1] = rng.integers(1000, 3200, n - 2)
        qm[b, :n] = 1
    a = np.zeros((3, 5), np.int64)
    am = np.zeros((3, 5), np.int64)
    for i, n in enumerate((5, 4, 3)):
        a[i, 0], a[i, n - 1] = 101, 102
        a[i, 1:n - 1] = rng.integers(1000, 3200, n - 2)
        am[i, :n] = 1
    return {"images": images, "question_ids": torch.from_numpy(q), "question_mask": torch.from_numpy(qm),
            "answer_ids": torch.from_numpy(a), "answer_mask": torch.from_numpy(am),
            "weights": torch.tensor([1.0, 0.5, 0.5]), "n": [1, 2], "alpha": 0.0, "train": True}
