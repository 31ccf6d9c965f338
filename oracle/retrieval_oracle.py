"""CPU restatement of the reference's similarity -> rank / top-k scoring (test infrastructure).

Follows ``/root/reference/cvap/module/decoder/loss_head.py``:
  * ``LossHead.infer``            :34-46    normalise unless told otherwise, stash
  * ``LossHead.retrieval_metrics`` :67-77   R@1/5/10/50, MED (lower median + 1), AVG (fp32 mean + 1)
  * ``LossHead.retrieval_eval``   :79-107   A->T min rank over the 5 captions, T->A rank, fp32 rank tensors
  * ``LossHead.report``           :109-244  N==M branch (:112-134), 1-vs-5 branch (:135-170), fallback (:171-174)
  * ``ClassificationHead.report`` :365-407  zero-shot: un-normalised ``audios @ text.t()``, argmax, label_map

The reference ranks with ``argsort(descending=True)`` + ``where``; on tie-free
rows the 0-based position of column j equals ``#{k : S[i,k] > S[i,j]}``, which
is what is computed here.  Where the reference formats a
0-d float32 *tensor* (int64 sum / python int -> float32) the same float32
arithmetic is reproduced so the strings are byte-identical.

Pinned against the reference's own output strings and rank vectors through
tests/golden/retrieval_*.npz|json (oracle/make_golden.py).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def normalize(x):
    x = np.asarray(x)
    return x / np.sqrt((x * x).sum(-1, keepdims=True))


def similarity(q, k, dtype=np.float32):
    return np.asarray(q, dtype) @ np.asarray(k, dtype).T


def rank_of(S, gt):
    """rank[i, c] = #{k : S[i,k] > S[i, gt[i,c]]}  (gt: (N,g) int)."""
    S = np.asarray(S)
    gt = np.asarray(gt)
    if gt.ndim == 1:
        gt = gt[:, None]
    ref = np.take_along_axis(S, gt, axis=1)                       # (N, g)
    return (S[:, None, :] > ref[:, :, None]).sum(-1).astype(np.int64)


def topk(S, k):
    """Indices/values of the k largest per row, descending, lower index first on ties."""
    S = np.asarray(S)
    order = np.lexsort((np.broadcast_to(np.arange(S.shape[1]), S.shape), -S), axis=1)[:, :k]
    return order.astype(np.int64), np.take_along_axis(S, order, axis=1)


def min_sim_margin(S64, gt):
    """Smallest |S[i,k] - S[i,gt]| over k != gt, in the fp64 ground truth."""
    S64 = np.asarray(S64, np.float64)
    gt = np.asarray(gt)
    if gt.ndim == 1:
        gt = gt[:, None]
    worst = np.inf
    rows = np.arange(S64.shape[0])
    for c in range(gt.shape[1]):
        d = np.abs(S64 - S64[rows, gt[:, c]][:, None])
        d[rows, gt[:, c]] = np.inf
        worst = min(worst, float(d.min()))
    return worst


def _pct_py(count, n):                      # python-float arithmetic (loss_head.py:70-73,121-122)
    return count / n * 100.0


def _pct_t(count, n):                       # int64 tensor / int -> float32 tensor, * 100. (:143-144)
    return f32(f32(count) / f32(n)) * f32(100.0)


def _mean1_t(r):                            # r.float().mean() + 1 (:146,160), float32 throughout
    r = np.asarray(r, dtype=f32)
    return f32(f32(r.sum(dtype=np.float64)) / f32(r.size)) + f32(1.0)


def retrieval_metrics(ranks, msg=""):
    """loss_head.py:67-77 on a float32 rank vector."""
    ranks = np.asarray(ranks, dtype=f32)
    n = ranks.shape[0]
    R1, R5, R10, R50 = (_pct_py(int((ranks < k).sum()), n) for k in (1, 5, 10, 50))
    MED = f32(np.sort(ranks)[(n - 1) // 2]) + f32(1.0)          # torch.median = lower median
    AVG = _mean1_t(ranks)
    return (f"{msg}: R@1 {R1:2.2f} R5 {R5:2.2f} R10 {R10:2.2f} R50 {R50:2.2f} "
            f"MED {float(MED):2.2f} AVG {float(AVG):2.2f}")


def retrieval_eval_from_ranks(r12, r21):
    """loss_head.py:79-107 given r12 (N,5) and r21 (5N,)."""
    return (retrieval_metrics(np.asarray(r12).min(-1), msg="A->T") + "\n" +
            retrieval_metrics(np.asarray(r21), msg="T->A"))


def report_strings_from_ranks(r12, r21, n1, n2):
    """The strings of LossHead.report() (:109-174, 241-244) without a gold file."""
    if n1 == n2:
        t12_1, t12_5 = _pct_py(int((r12 < 1).sum()), n1), _pct_py(int((r12 < 5).sum()), n1)
        t21_1, t21_5 = _pct_py(int((r21 < 1).sum()), n1), _pct_py(int((r21 < 5).sum()), n1)
        p_12 = f"I->A: t1 = {t12_1:2.2f} t5 = {t12_5:2.2f}"
        p_21 = f"A->I: t1 = {t21_1:2.2f} t5 = {t21_5:2.2f}"
        ref = ""
    elif n1 * 5 == n2:
        r12 = np.asarray(r12).reshape(-1, 5)
        t12_1 = _pct_t(int((r12 < 1).sum()), 1 * r12.shape[0])
        t12_5 = _pct_t(int((r12 < 5).sum()), 5 * r12.shape[0])
        mean12 = _mean1_t(r12.min(-1))
        p_12 = f"A->T: t1 = {float(t12_1):2.2f} t5 = {float(t12_5):2.2f} mR = {float(mean12):2.2f}"
        t21_1 = _pct_py(int((r21 < 1).sum()), r21.shape[0])
        t21_5 = _pct_py(int((r21 < 5).sum()), r21.shape[0])
        mean21 = _mean1_t(r21)
        p_21 = f"T->A: t1 = {t21_1:2.2f} t5 = {t21_5:2.2f} mR = {float(mean21):2.2f}"
        ref = "\nREFERENCE\n" + retrieval_eval_from_ranks(r12, r21)
    else:
        raise ValueError("shape relation handled by the caller (loss_head.py:171-174)")
    return f"{p_12} {p_21} @ {n1}{ref}"


def report(x1s, x2s, dtype=np.float32):
    """LossHead.report() for already-stashed (normalised) features, no gold file."""
    x1s, x2s = np.asarray(x1s, dtype), np.asarray(x2s, dtype)
    n1, n2 = x1s.shape[0], x2s.shape[0]
    if n1 == n2:
        gt = np.arange(n1)
        r12 = rank_of(similarity(x1s, x2s, dtype), gt)[:, 0]
        r21 = rank_of(similarity(x2s, x1s, dtype), gt)[:, 0]
    elif n1 * 5 == n2:
        gt12 = np.arange(n2).reshape(n1, 5)
        r12 = rank_of(similarity(x1s, x2s, dtype), gt12)
        r21 = rank_of(similarity(x2s, x1s, dtype), np.arange(n2) // 5)[:, 0]
    else:
        return f"torch.Size([{n1}, {x1s.shape[1]}])xtorch.Size([{n2}, {x2s.shape[1]}]) - @ {n1}", None, None
    return report_strings_from_ranks(r12, r21, n1, n2), r12, r21


def zero_shot_report(audios, text, labels, label_map=None, dtype=np.float32):
    """ClassificationHead.report(text=...) (:365-407): argmax of the UN-normalised similarity."""
    S = similarity(audios, text, dtype)
    pred = S.argmax(-1)
    if isinstance(label_map, dict):
        pred = np.asarray([label_map[int(p)] for p in pred])
    labels = np.asarray(labels)
    n = labels.shape[0]
    precision = _pct_t(int((pred == labels).sum()), n)
    return f"A->T: p1 = {float(precision):2.2f} @ {n}", pred.astype(np.int64)
