"""Seeded synthetic VQA batches: there are no datasets on the box, so the hot
path is fed tensors of the real shapes and value ranges.  Used by tests, bench.py and the
``--climb_data_dir synthetic`` mode of the round loop.
"""
from __future__ import annotations

from typing import Dict, Iterator, Optional

import torch

SCORES = (0.3, 0.6, 0.9, 1.0)          # VQA soft scores, reference src/utils/vqa_utils.py:21-31


def make_vilt_batch(batch_size: int, text_len: int = 40, image_size: int = 384, num_labels: int = 100,
                    seed: int = 0, client: Optional[int] = None, pin: bool = False) -> Dict:
    """Pre-encoded ViLT inputs + soft VQA targets on the HOST (pinned when ``pin``).

    ``client`` selects a heterogeneous synthetic domain (BASELINE config 3): pixel mean shift
    0.25 * (c - 3.5), text-id sub-range [1000 + 3500 c, 1000 + 3500 (c + 1)), label prior
    concentrated on labels [12 c, 12 c + 24).
    """
    g = torch.Generator().manual_seed(seed)
    lo, hi = 1000, 30000
    shift = 0.0
    label_lo, label_hi = 0, num_labels
    if client is not None:
        lo, hi = 1000 + 3500 * client, 1000 + 3500 * (client + 1)
        shift = 0.25 * (client - 3.5)
        label_lo = (12 * client) % num_labels
        label_hi = min(num_labels, label_lo + 24)
    enc = {
        "input_ids": torch.randint(lo, hi, (batch_size, text_len), generator=g, dtype=torch.int64),
        "attention_mask": torch.ones(batch_size, text_len, dtype=torch.int64),
        "token_type_ids": torch.zeros(batch_size, text_len, dtype=torch.int64),
        "pixel_values": torch.randn(batch_size, 3, image_size, image_size, generator=g) + shift,
        "pixel_mask": torch.ones(batch_size, image_size, image_size, dtype=torch.int64),
    }
    target = torch.zeros(batch_size, num_labels)
    for i in range(batch_size):
        k = int(torch.randint(1, 4, (1,), generator=g))
        labels = label_lo + torch.randperm(label_hi - label_lo, generator=g)[:k]
        scores = torch.tensor(SCORES)[torch.randint(0, 4, (k,), generator=g)]
        target[i, labels] = scores
    batch = {"encodings": enc, "target_scores": target}
    if pin and torch.cuda.is_available():
        batch = {"encodings": {k: v.pin_memory() for k, v in enc.items()}, "target_scores": target.pin_memory()}
    return batch


def to_device(batch: Dict, device, non_blocking: bool = True, image_dtype=torch.bfloat16) -> Dict:
    """Host -> device copy of one batch (the only H2D traffic of a train step)."""
    enc = {}
    for k, v in batch["encodings"].items():
        v = v.to(device, non_blocking=non_blocking)
        if k == "pixel_values" and image_dtype is not None:
            v = v.to(image_dtype)
        enc[k] = v
    enc["dense_masks"] = True      # synthetic masks are all ones: the dense fast path applies
    return {"encodings": enc, "target_scores": batch["target_scores"].to(device, non_blocking=non_blocking)}


class SyntheticVQALoader:
    """Deterministic stream of pre-encoded batches for one client (stands in for the reference's
    per-client DataLoader, train_vqa_crossvqa.py:129-230)."""

    def __init__(self, num_batches: int, batch_size: int, device, text_len=40, image_size=384, num_labels=100,
                 seed=0, client=None, image_dtype=torch.bfloat16):
        self.args = dict(batch_size=batch_size, text_len=text_len, image_size=image_size, num_labels=num_labels,
                         client=client)
        self.num_batches, self.seed, self.device, self.image_dtype = num_batches, seed, device, image_dtype

    def __len__(self):
        return self.num_batches

    def __iter__(self) -> Iterator[Dict]:
        for i in range(self.num_batches):
            host = make_vilt_batch(seed=self.seed * 100003 + i, pin=torch.device(self.device).type == "cuda",
                                   **self.args)
            yield to_device(host, self.device, image_dtype=self.image_dtype)


# ------------------------------------------------------------------------------------------------ ALBEF
CLS_ID, SEP_ID, PAD_ID = 101, 102, 0        # bert-base-uncased special tokens


def make_albef_batch(batch_size: int, image_size: int = 384, seed: int = 0, client: Optional[int] = None,
                     q_len: int = 25, a_len: int = 6, vocab: int = 30522, pin: bool = False) -> Dict:
    """Pre-tokenised ALBEF VQA training batch on the HOST (the collate contract of reference
    vqa_dataset_crossvqa.py:456-471): images N(0, 1); questions [CLS] ids... [SEP] padded to ``q_len``; k_b in
    {1, 2, 3} answers per sample ([CLS] a... [SEP], 3..``a_len`` tokens, padded) with weights 1 / k_b.  k_b cycles
    deterministically so that the number of answer sequences is the same for every batch of a size (static shapes
    for CUDA-graph replay)."""
    g = torch.Generator().manual_seed(seed)
    lo, hi, shift = 1000, vocab, 0.0
    if client is not None:
        span = (vocab - 1000) // 8
        lo, hi = 1000 + span * client, 1000 + span * (client + 1)
        shift = 0.25 * (client - 3.5)
    images = torch.randn(batch_size, 3, image_size, image_size, generator=g) + shift
    q_ids = torch.full((batch_size, q_len), PAD_ID, dtype=torch.int64)
    q_mask = torch.zeros(batch_size, q_len, dtype=torch.int64)
    for b in range(batch_size):
        n = int(torch.randint(6, q_len + 1, (1,), generator=g))
        q_ids[b, 0], q_ids[b, n - 1] = CLS_ID, SEP_ID
        q_ids[b, 1:n - 1] = torch.randint(lo, hi, (n - 2,), generator=g)
        q_mask[b, :n] = 1
    ks = [1 + ((b + seed) % batch_size) % 3 for b in range(batch_size)]     # a rotation of one list: constant sum
    n_seq = sum(ks)
    a_ids = torch.full((n_seq, a_len), PAD_ID, dtype=torch.int64)
    a_mask = torch.zeros(n_seq, a_len, dtype=torch.int64)
    weights = torch.empty(n_seq)
    i = 0
    for b, k in enumerate(ks):
        for _ in range(k):
            n = int(torch.randint(3, a_len + 1, (1,), generator=g))
            a_ids[i, 0], a_ids[i, n - 1] = CLS_ID, SEP_ID
            a_ids[i, 1:n - 1] = torch.randint(lo, hi, (n - 2,), generator=g)
            a_mask[i, :n] = 1
            weights[i] = 1.0 / k
            i += 1
    batch = {"images": images, "question_ids": q_ids, "question_mask": q_mask, "answer_ids": a_ids,
             "answer_mask": a_mask, "weights": weights, "n": ks, "alpha": 0.0, "train": True,
             # question of each answer sequence (albef_model.py:92-98 builds it per forward from ``n``)
             "answer_index": torch.repeat_interleave(torch.arange(batch_size), torch.tensor(ks))}
    if pin and torch.cuda.is_available():
        batch = {k: (v.pin_memory() if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
    return batch


def albef_to_device(batch: Dict, device, image_dtype=torch.bfloat16, non_blocking: bool = True) -> Dict:
    out = {}
    for k, v in batch.items():
        if isinstance(v, torch.Tensor):
            v = v.to(device, non_blocking=non_blocking)
            if k == "images" and image_dtype is not None:
                v = v.to(image_dtype)
        out[k] = v
    return out


def make_albef_eval_set(num_batches: int, batch_size: int, image_size: int, seed: int, n_answers: int = 96,
                        a_len: int = 6, q_len: int = 25, vocab: int = 30522):
    """(answer_list_ids, answer_list_mask, [(images, question_ids, question_mask, gts)]) for the rank_answer
    evaluation (reference task_trainer.py:157-205: top-k over a fixed answer list, exact match against gts)."""
    g = torch.Generator().manual_seed(seed)
    ans = torch.full((n_answers, a_len), PAD_ID, dtype=torch.int64)
    mask = torch.zeros(n_answers, a_len, dtype=torch.int64)
    for i in range(n_answers):
        n = int(torch.randint(3, a_len + 1, (1,), generator=g))
        ans[i, 0], ans[i, n - 1] = CLS_ID, SEP_ID
        ans[i, 1:n - 1] = torch.randint(1000, vocab, (n - 2,), generator=g)
        mask[i, :n] = 1
    batches = []
    for j in range(num_batches):
        tr = make_albef_batch(batch_size, image_size, seed=seed * 977 + j, q_len=q_len, vocab=vocab)
        gts = torch.randint(0, n_answers, (batch_size, 1), generator=g)
        batches.append((tr["images"], tr["question_ids"], tr["question_mask"], gts))
    return ans, mask, batches


class SyntheticAlbefLoader:
    """Deterministic stream of pre-tokenised ALBEF batches for one client."""

    def __init__(self, num_batches: int, batch_size: int, device, image_size=384, seed=0, client=None,
                 image_dtype=torch.bfloat16):
        self.num_batches, self.batch_size, self.device = num_batches, batch_size, device
        self.image_size, self.seed, self.client, self.image_dtype = image_size, seed, client, image_dtype

    def __len__(self):
        return self.num_batches

    def __iter__(self) -> Iterator[Dict]:
        for i in range(self.num_batches):
            host = make_albef_batch(self.batch_size, self.image_size, seed=self.seed * 100003 + i, client=self.client,
                                    pin=torch.device(self.device).type == "cuda")
            yield albef_to_device(host, self.device, self.image_dtype)
