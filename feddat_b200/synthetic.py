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
