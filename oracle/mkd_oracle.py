"""Oracle: the MKD loss head (reference src/train/visionlanguage_tasks/task_trainer.py).  TEST ONLY."""
from __future__ import annotations

import numpy as np


def _log_softmax(x, axis):
    m = x.max(axis=axis, keepdims=True)
    z = x - m
    return z - np.log(np.exp(z).sum(axis=axis, keepdims=True))


def kl_loss(output, target, temp=3.0, with_grad=False):
    """task_trainer.py:506-516.  softmax over the LAST dim when it is > 3000 wide (:507-509), else
    over dim=1 (:510-512); F.kl_div(p_log, q, 'batchmean') divides by output.shape[0] (:514); times
    temp**2 (:515)."""
    out = np.asarray(output, np.float64)
    tgt = np.asarray(target, np.float64)
    axis = -1 if out.shape[-1] > 3000 else 1
    p_log = _log_softmax(out / temp, axis)
    q_log = _log_softmax(tgt / temp, axis)
    q = np.exp(q_log)
    point = np.where(q > 0, q * (q_log - p_log), 0.0)
    loss = point.sum() / out.shape[0] * temp ** 2
    if not with_grad:
        return loss
    grad = (np.exp(p_log) - q) * temp / out.shape[0]
    return loss, grad


def bce_with_logits_times_c(logits, target, with_grad=False):
    """task_trainer.py:299 -- nn.BCEWithLogitsLoss(reduction='mean')(logits, target) * target.shape[1]
    (criterion built at train_vqa_crossvqa.py:237)."""
    x = np.asarray(logits, np.float64)
    t = np.asarray(target, np.float64)
    l = np.maximum(x, 0) - x * t + np.log1p(np.exp(-np.abs(x)))
    loss = l.mean() * t.shape[1]
    if not with_grad:
        return loss
    sig = 1.0 / (1.0 + np.exp(-x))
    return loss, (sig - t) / x.shape[0]


def mkd_total(logits, teacher, target, temp=3.0):
    """task_trainer.py:299-301 / :319-321:  L = (task + kl_loss(logits, teacher.detach())) / 2.
    Returns (L, kl, task, dL/dlogits)."""
    kl, gkl = kl_loss(logits, teacher, temp, with_grad=True)
    task, gtask = bce_with_logits_times_c(logits, target, with_grad=True)
    return (task + kl) / 2.0, kl, task, (gkl + gtask) / 2.0


def albef_answer_loss(prediction_scores, labels, weights, batch_size, with_grad=False):
    """The ALBEF task loss: BertLMHeadModel.forward with reduction='none'
    (src/modeling/models/xbert.py:1287-1297: scores[:, :-1] against labels[:, 1:], CrossEntropyLoss
    ignore_index = -100, per-sequence sums) weighted and normalised as ALBEF.forward does
    (src/modeling/models/albef_model.py:142-143: (weights * loss).sum() / image.size(0))."""
    s = np.asarray(prediction_scores, np.float64)
    lab = np.asarray(labels)
    w = np.asarray(weights, np.float64)
    shifted, tgt = s[:, :-1, :], lab[:, 1:]
    logp = _log_softmax(shifted, -1)
    valid = tgt != -100
    safe = np.where(valid, tgt, 0)
    tok = -np.take_along_axis(logp, safe[..., None], axis=-1)[..., 0] * valid
    loss = float((w * tok.sum(1)).sum() / batch_size)
    if not with_grad:
        return loss
    g = np.exp(logp)
    np.put_along_axis(g, safe[..., None], np.take_along_axis(g, safe[..., None], axis=-1) - 1.0, axis=-1)
    g = g * valid[..., None] * (w[:, None, None] / batch_size)
    grad = np.zeros_like(s)
    grad[:, :-1, :] = g
    return loss, grad


def mkd_ce_total(prediction_scores, teacher_shifted, labels, weights, batch_size, temp=3.0):
    """task_trainer.py:296-301 on the ALBEF branch: L = (answer_loss + kl_loss(logits[:, :-1], teacher)) / 2 with
    logits[:, :-1] what ALBEF.forward returns (albef_model.py:145).  Returns (L, kl, task, dL/dscores)."""
    s = np.asarray(prediction_scores, np.float64)
    kl, gkl = kl_loss(s[:, :-1, :], teacher_shifted, temp, with_grad=True)
    task, gtask = albef_answer_loss(s, labels, weights, batch_size, with_grad=True)
    grad = gtask / 2.0
    grad[:, :-1, :] += gkl / 2.0
    return (task + kl) / 2.0, kl, task, grad
