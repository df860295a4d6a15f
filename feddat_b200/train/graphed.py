"""CUDA-graph replay of ``TaskTrainer.train_step`` (dat mode).

One train step is ~1000 kernel launches (three ViLT forwards, two backwards, two fused-AdamW steps);
launched eagerly from Python the GPU idles a third of the time.  The whole step -- including the
autograd backward, the sm_100a DAT / MKD kernels (enqueued through the C ABI on the capturing stream)
and both optimizer steps -- is captured once and replayed per batch.

What stays outside the graph, per step: one copy of the batch into static tensors (``prefetch`` stages the
NEXT batch host->device on a side stream while the current step computes) and one tiny H2D
copy of the three learning-rate values the step needs (the reference steps its scheduler twice per
batch: task_trainer.py:303-308, 323-328), computed on the host by the scheduler's own lambda.
Semantics are those of ``train_step`` (same call order, same detach points, grads set to None).
"""
from __future__ import annotations

from typing import Dict

import torch


class _ReplayScheduler:
    """Stands in for the LR scheduler inside the captured region: each ``step()`` copies the next
    pre-staged learning rate into the optimizer's (tensor) lr."""

    def __init__(self, optimizer, lr_buf):
        self.optimizer, self.lr_buf, self.i = optimizer, lr_buf, 0

    def load(self, idx):
        for g in self.optimizer.param_groups:
            g["lr"].copy_(self.lr_buf[idx])

    def step(self):
        self.i += 1
        self.load(self.i)


class GraphedTrainStep:
    def __init__(self, trainer, wrapped_model, optimizer, scheduler, example_batch: Dict, warmup: int = 2):
        for g in optimizer.param_groups:
            if not (isinstance(g["lr"], torch.Tensor) and g["lr"].is_cuda and g.get("capturable", False)):
                raise ValueError("GraphedTrainStep needs an optimizer built by TaskTrainer.create_optimizer on "
                                 "CUDA (capturable fused AdamW with a tensor lr)")
        self.trainer, self.model, self.opt, self.sched = trainer, wrapped_model, optimizer, scheduler
        dev = example_batch["target_scores"].device
        self.static = {"encodings": {k: (v.clone() if isinstance(v, torch.Tensor) else v)
                                     for k, v in example_batch["encodings"].items()},
                       "target_scores": example_batch["target_scores"].clone()}
        self.lr_buf = torch.zeros(3, device=dev, dtype=torch.float32)
        self.base_lrs = [float(b) for b in scheduler.base_lrs]
        self.lmbda = scheduler.lr_lambdas[0]
        # host->device staging for prefetch(): filled on a copy stream while the previous step runs
        self._copy_stream = torch.cuda.Stream(device=dev)
        self._staged = None
        self._staged_ready = torch.cuda.Event()
        self._staged_free = torch.cuda.Event()
        self._staged_free.record()
        self.graph = None
        self.loss = None
        self.launches_per_step = 0
        # eager warm-up steps on a side stream (real training steps: state advances exactly as eager)
        self._warm = warmup

    # ------------------------------------------------------------------
    def _stage_lrs(self):
        e = self.sched.last_epoch
        vals = [self.base_lrs[0] * self.lmbda(e + i) for i in range(3)]
        host = torch.tensor(vals, dtype=torch.float32).pin_memory()
        self.lr_buf.copy_(host, non_blocking=True)

    def _advance_scheduler(self):
        self.sched.last_epoch += 2
        self.sched._last_lr = [b * self.lmbda(self.sched.last_epoch) for b in self.base_lrs]

    def _used_keys(self, enc):
        """Tensors the captured forward actually reads: the dense fast path (all-ones masks) never
        touches attention_mask / pixel_mask, so they are not copied (pixel_mask alone is 38 MB of int64
        per 32 x 384 x 384 batch)."""
        dense = self.static["encodings"].get("dense_masks", False)      # host batches carry no flag
        if dense and self.model.module.vilt_encoder.dense_fast_path:
            return [k for k in ("input_ids", "token_type_ids", "pixel_values") if isinstance(enc.get(k), torch.Tensor)]
        return [k for k, v in enc.items() if isinstance(v, torch.Tensor)]

    def _load_batch(self, batch):
        enc = batch["encodings"]
        for k in self._used_keys(enc):
            self.static["encodings"][k].copy_(enc[k], non_blocking=True)
        self.static["target_scores"].copy_(batch["target_scores"], non_blocking=True)

    def prefetch(self, batch: Dict) -> None:
        """Start the host->device copy of the NEXT batch on the copy stream (pinned host tensors);
        the next ``__call__()`` without a batch consumes it with a device-to-device copy."""
        enc = batch["encodings"]
        keys = self._used_keys(enc)
        if self._staged is None:
            # staged in the HOST dtype (fp32 pixels): a dtype-converting H2D copy would convert on the CPU
            dev = self.static["target_scores"].device
            self._staged = {k: torch.empty(enc[k].shape, dtype=enc[k].dtype, device=dev) for k in keys}
            self._staged["target_scores"] = torch.empty_like(self.static["target_scores"])
        self._copy_stream.wait_event(self._staged_free)      # the previous consumer is done with it
        with torch.cuda.stream(self._copy_stream):
            for k in keys:
                self._staged[k].copy_(enc[k], non_blocking=True)
            self._staged["target_scores"].copy_(batch["target_scores"], non_blocking=True)
            self._staged_ready.record(self._copy_stream)
        self._staged_keys = keys

    def _load_staged(self):
        cur = torch.cuda.current_stream()
        cur.wait_event(self._staged_ready)
        for k in self._staged_keys:
            self.static["encodings"][k].copy_(self._staged[k], non_blocking=True)
        self.static["target_scores"].copy_(self._staged["target_scores"], non_blocking=True)
        self._staged_free.record(cur)

    def accepts(self, batch) -> bool:
        """True when ``batch`` has the captured shapes (a ragged last batch must run eagerly)."""
        if batch["target_scores"].shape != self.static["target_scores"].shape:
            return False
        return all(not isinstance(v, torch.Tensor) or v.shape == self.static["encodings"][k].shape
                   for k, v in batch["encodings"].items())

    def _clear_caches(self):
        self.model.module.vilt_encoder._embed_cache = None

    def _body(self):
        rs = _ReplayScheduler(self.opt, self.lr_buf)
        rs.load(0)
        return self.trainer.train_step(self.model, 0, self.static, self.opt, rs)

    def _capture(self):
        from .. import ops
        self._clear_caches()
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        n0 = ops.launch_count
        with torch.cuda.graph(self.graph):
            self.loss = self._body()
        self.launches_per_step = ops.launch_count - n0
        self._clear_caches()          # the cached embedding now lives in the graph's private pool: never reuse eagerly

    # ------------------------------------------------------------------
    def __call__(self, batch: Dict = None) -> torch.Tensor:
        """Runs one train step on ``batch`` (host-pinned or device tensors), or on the batch staged by
        ``prefetch`` when ``batch`` is None; returns the task loss (device scalar, valid until the
        next call)."""
        if batch is None:
            self._load_staged()
        else:
            self._load_batch(batch)
        if self._warm > 0:
            self._warm -= 1
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                loss = self.trainer.train_step(self.model, 0, self.static, self.opt, self.sched)
            torch.cuda.current_stream().wait_stream(s)
            return loss
        self._stage_lrs()
        if self.graph is None:
            self._capture()
        self.graph.replay()
        self._advance_scheduler()
        from .. import ops
        ops.launch_count += self.launches_per_step
        return self.loss


class GraphedDictStep(GraphedTrainStep):
    """The same capture / replay for learners whose batch is a FLAT dict of tensors and Python values (the ALBEF
    path: images, question / answer ids and masks, weights, answer_index; ``n`` / ``alpha`` / ``train`` are Python
    values and must not change between batches).  No prefetch staging: ``__call__(batch)`` copies the batch's tensors
    into the static ones (host-pinned or device sources)."""

    def __init__(self, trainer, wrapped_model, optimizer, scheduler, example_batch: Dict, warmup: int = 2,
                 unused=("n",)):
        """``unused``: Python-valued batch entries the captured step does not read (ALBEF's per-sample answer counts
        ``n`` are superseded by the ``answer_index`` tensor) and which may therefore differ between batches."""
        self._unused = tuple(unused)
        for g in optimizer.param_groups:
            if not (isinstance(g["lr"], torch.Tensor) and g["lr"].is_cuda and g.get("capturable", False)):
                raise ValueError("GraphedDictStep needs an optimizer built by TaskTrainer.create_optimizer on CUDA")
        self.trainer, self.model, self.opt, self.sched = trainer, wrapped_model, optimizer, scheduler
        self.static = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in example_batch.items()}
        dev = next(v.device for v in self.static.values() if isinstance(v, torch.Tensor))
        if dev.type != "cuda":
            raise ValueError("GraphedDictStep: the example batch must live on the GPU")
        self.lr_buf = torch.zeros(3, device=dev, dtype=torch.float32)
        self.base_lrs = [float(b) for b in scheduler.base_lrs]
        self.lmbda = scheduler.lr_lambdas[0]
        self.graph = None
        self.loss = None
        self.launches_per_step = 0
        self._warm = warmup

    def accepts(self, batch) -> bool:
        return all((isinstance(v, torch.Tensor) and v.shape == self.static[k].shape)
                   if isinstance(self.static.get(k), torch.Tensor) else (k in self._unused or v == self.static.get(k))
                   for k, v in batch.items())

    def _load_batch(self, batch):
        for k, v in batch.items():
            if isinstance(v, torch.Tensor):
                self.static[k].copy_(v, non_blocking=True)
            elif k not in self._unused and v != self.static[k]:
                raise ValueError(f"GraphedDictStep: batch[{k!r}] = {v!r} differs from the captured {self.static[k]!r}")

    def _load_staged(self):
        raise RuntimeError("GraphedDictStep has no prefetch stage: pass the batch to __call__")

    def prefetch(self, batch):
        raise RuntimeError("GraphedDictStep has no prefetch stage: pass the batch to __call__")

    def _clear_caches(self):
        pass
