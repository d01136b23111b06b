"""Evaluation metrics computed where the predictions live.

The reference moves labels and predictions of every batch to Python lists and hands them to
``sklearn.metrics.roc_auc_score`` / ``log_loss`` (trainers/ctr_trainer.py:99-111, 113-152).  On a CUDA device the
trainer keeps both on the device and evaluates the same two definitions there in float64; only the scalar results
cross to the host.  Both functions work on any device (the CPU tests hold them against sklearn).

* ``binary_auc``: area under the ROC curve with ties at half weight = the Mann-Whitney statistic on mid-ranks, which is
  what sklearn's trapezoidal rule over the distinct thresholds evaluates to.
* ``binary_logloss``: sklearn.metrics.log_loss for labels {0, 1}: probabilities clipped to [eps, 1 - eps] with
  eps = float64 machine epsilon (what sklearn uses for the float64 arrays it builds from Python lists), mean of
  -[y log p + (1 - y) log(1 - p)].

Error behaviour follows sklearn: a single class in ``y_true`` raises ``ValueError``.
"""
from __future__ import annotations

import torch

_EPS64 = float(torch.finfo(torch.float64).eps)


def _check_binary(y: torch.Tensor) -> None:
    if y.numel() == 0:
        raise ValueError("empty input")
    if not bool(((y == 0) | (y == 1)).all()):
        raise ValueError("binary metrics need labels in {0, 1}")


def binary_auc(y_true: torch.Tensor, y_score: torch.Tensor) -> float:
    y = y_true.reshape(-1).to(torch.float64)
    s = y_score.reshape(-1).to(device=y.device, dtype=torch.float64)
    _check_binary(y)
    n = y.numel()
    n_pos = float(y.sum())
    n_neg = n - n_pos
    if n_pos == 0 or n_neg == 0:
        raise ValueError("Only one class present in y_true. ROC AUC score is not defined in that case.")
    _uniq, inv, counts = torch.unique(s, sorted=True, return_inverse=True, return_counts=True)
    cum = counts.cumsum(0).to(torch.float64)
    midrank = cum - (counts.to(torch.float64) - 1.0) * 0.5          # 1-based mean rank of every distinct score
    rank_sum_pos = float((midrank[inv] * y).sum())
    return (rank_sum_pos - n_pos * (n_pos + 1.0) * 0.5) / (n_pos * n_neg)


def binary_logloss(y_true: torch.Tensor, y_pred: torch.Tensor) -> float:
    y = y_true.reshape(-1).to(torch.float64)
    p = y_pred.reshape(-1).to(device=y.device, dtype=torch.float64)
    _check_binary(y)
    if float(y.min()) == float(y.max()):
        raise ValueError("y_true contains only one label. Please provide the true labels explicitly through the labels argument.")
    p = p.clamp(_EPS64, 1.0 - _EPS64)
    return float(-(y * torch.log(p) + (1.0 - y) * torch.log1p(-p)).mean())
