"""CPU restatement of the AudioSet zero-shot / multi-label scoring that follows the similarity product
(TEST INFRASTRUCTURE for a kernel that does not exist yet -- built first so that the
GPU scorer has a checker the day it is written).

Follows ``/root/reference/cvap/module/decoder/loss_more.py``:
  * ``BCELossHead.zero_shot``  :77-84   normalise audios and label-text embeddings, ``x1s = audios @ text.t()``
  * ``BCELossHead.report``     :86-131  micro / macro / weighted AP over the (N, C) score matrix, then per class
                                        AP, ROC-AUC and the MIDDLE point of the precision-recall curve; report string.

The arithmetic lives in a third-party dependency that is not under ``/root/reference``: scikit-learn, pinned to
``scikit_learn==1.0.1`` (``requirements.txt:15``).  Restated here from its published algorithm
(``sklearn/metrics/_ranking.py``):
  * ``_binary_clf_curve``: stable sort by descending score, cumulative tp / fp at the LAST index of every distinct score;
  * ``average_precision_score`` (binary): ``sum_n (R_n - R_{n-1}) P_n`` over the distinct-threshold PR curve;
    micro = the same on the raveled (N*C) arrays, macro = mean over classes, weighted = mean weighted by class support;
  * ``roc_auc_score`` (binary): trapezoid rule over (fpr, tpr) with the (0, 0) origin prepended;
  * ``precision_recall_curve``: 1.0.1 TRUNCATES the curve at the first threshold that reaches full recall, releases
    >= 1.1 do not.  The middle point ``p[len(p)//2], r[len(p)//2]`` the reference reports (mP / mR) therefore depends on
    the scikit-learn version; ``truncate=True`` (default) is the pinned 1.0.1 behaviour.

Pinning: AP (binary / micro / macro / weighted), AUC and the un-truncated PR curve are checked against the scikit-learn
installed in the build container (1.9) in tests/test_oracle_map.py; the truncated curve against the known-answer vector
of the 1.0.1 documentation.  The report string with 1.0.1's mP / mR is therefore pinned by restatement + known answer only.
"""
from __future__ import annotations

import math

import numpy as np


def binary_clf_curve(y_true, y_score):
    """(fps, tps, thresholds) at every distinct score, descending (sklearn _binary_clf_curve, pos_label = 1)."""
    y_true = np.asarray(y_true) == 1
    y_score = np.asarray(y_score, dtype=np.float64)
    order = np.argsort(y_score, kind="mergesort")[::-1]
    y_score, y_true = y_score[order], y_true[order]
    distinct = np.where(np.diff(y_score))[0]
    idx = np.r_[distinct, y_true.size - 1]
    tps = np.cumsum(y_true, dtype=np.float64)[idx]
    fps = 1 + idx - tps
    return fps, tps, y_score[idx]


def precision_recall_curve(y_true, y_score, truncate=True):
    fps, tps, thr = binary_clf_curve(y_true, y_score)
    ps = tps + fps
    precision = np.divide(tps, ps, out=np.zeros_like(tps), where=ps != 0)
    recall = np.ones_like(tps) if tps[-1] == 0 else tps / tps[-1]
    if truncate:                                  # scikit-learn 1.0.1: stop once full recall is attained
        last = int(tps.searchsorted(tps[-1]))
        sl = slice(last, None, -1)
    else:
        sl = slice(None, None, -1)
    return np.r_[precision[sl], 1.0], np.r_[recall[sl], 0.0], thr[sl]


def average_precision(y_true, y_score):
    """Binary AP = sum_n (R_n - R_{n-1}) P_n; NaN when the class has no positive (as sklearn: 0/0)."""
    fps, tps, _ = binary_clf_curve(y_true, y_score)
    if tps[-1] == 0:
        return float("nan")
    precision = tps / (tps + fps)
    recall = tps / tps[-1]
    return float(np.sum(np.diff(np.r_[0.0, recall]) * precision))


def roc_auc(y_true, y_score):
    """Binary ROC-AUC (trapezoid); raises ValueError when only one class is present (sklearn does)."""
    fps, tps, _ = binary_clf_curve(y_true, y_score)
    if tps[-1] == 0 or fps[-1] == 0:
        raise ValueError("Only one class present in y_true. ROC AUC score is not defined in that case.")
    fpr, tpr = np.r_[0.0, fps] / fps[-1], np.r_[0.0, tps] / tps[-1]
    return float(np.sum(np.diff(fpr) * (tpr[1:] + tpr[:-1]) * 0.5))


def average_precision_multilabel(Y, S, average):
    Y, S = np.asarray(Y), np.asarray(S)
    if average == "micro":
        return average_precision(Y.ravel(), S.ravel())
    per = np.array([average_precision(Y[:, k], S[:, k]) for k in range(Y.shape[1])])
    if average == "macro":
        return float(np.mean(per))
    if average == "weighted":
        w = Y.sum(0).astype(np.float64)
        return float(np.sum(per * w) / w.sum()) if w.sum() > 0 else float("nan")
    raise ValueError(average)


def normalize(x):
    x = np.asarray(x)
    return x / np.sqrt((x * x).sum(-1, keepdims=True))


def zero_shot_scores(audios, text, normalized=False, dtype=np.float32):
    """``BCELossHead.zero_shot`` :77-83 -- both sides normalised unless the head says they already are."""
    a, t = np.asarray(audios, dtype), np.asarray(text, dtype)
    if not normalized:
        a, t = normalize(a), normalize(t)
    return a @ t.T


def report(x1s, x2s, truncate=True):
    """``BCELossHead.report`` :92-130: the string, plus the per-class lists for finer checks."""
    x1s, x2s = np.asarray(x1s), np.asarray(x2s)
    nsample, nlabel = x1s.shape[:2]
    ap_micro = average_precision_multilabel(x2s, x1s, "micro")
    ap_macro = average_precision_multilabel(x2s, x1s, "macro")
    ap_weighted = average_precision_multilabel(x2s, x1s, "weighted")
    has_err = False
    ap_list, auc_list, precisions, recalls = [], [], [], []
    for k in range(nlabel):
        y_true, y_score = x2s[:, k], x1s[:, k]
        ap = average_precision(y_true, y_score)
        if math.isnan(ap):
            ap, has_err = 0.0, True
        try:
            auc = roc_auc(y_true, y_score)
        except ValueError:
            auc, has_err = 0.0, True
        p, r, _ = precision_recall_curve(y_true, y_score, truncate=truncate)
        mid = len(p) // 2
        ap_list.append(ap)
        auc_list.append(auc)
        precisions.append(p[mid])
        recalls.append(r[mid])
    mean_ap, mean_auc = np.mean(ap_list) * 100.0, np.mean(auc_list) * 100.0
    mean_p, mean_r = np.mean(precisions) * 100.0, np.mean(recalls) * 100.0
    text = f"Err({has_err}) mAP = {mean_ap:2.2f} mAUC = {mean_auc:2.2f} mP = {mean_p:2.2f} mR = {mean_r:2.2f}"
    common = f"Mac-AP = {ap_macro:2.2f} Mic-AP = {ap_micro:2.2f} wAP = {ap_weighted:2.2f}"
    return f"{common} {text} @ {nsample}", dict(ap=ap_list, auc=auc_list, p_mid=precisions, r_mid=recalls)
