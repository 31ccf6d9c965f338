"""The multi-label scoring oracle (oracle/map_oracle.py) against scikit-learn -- the library the
reference calls (loss_more.py:96-112) -- and against the known-answer vector of the release the reference pins (1.0.1)."""
import warnings

import numpy as np
import pytest

from oracle import map_oracle as mo

metrics = pytest.importorskip("sklearn.metrics")


def _multilabel(n, c, seed, ties=False):
    rng = np.random.default_rng(seed)
    Y = (rng.random((n, c)) < 0.08).astype(np.int64)
    S = rng.standard_normal((n, c)).astype(np.float32) + 1.5 * Y
    if ties:
        S = np.round(S * 4) / 4          # many equal scores: the distinct-threshold logic matters
    return Y, S


@pytest.mark.parametrize("ties", [False, True])
def test_ap_and_auc_match_sklearn(ties):
    Y, S = _multilabel(400, 12, 0, ties)
    for k in range(Y.shape[1]):
        assert mo.average_precision(Y[:, k], S[:, k]) == pytest.approx(metrics.average_precision_score(Y[:, k], S[:, k]), rel=1e-12)
        assert mo.roc_auc(Y[:, k], S[:, k]) == pytest.approx(metrics.roc_auc_score(Y[:, k], S[:, k]), rel=1e-12)
    for avg in ("micro", "macro", "weighted"):
        assert mo.average_precision_multilabel(Y, S, avg) == pytest.approx(metrics.average_precision_score(Y, S, average=avg), rel=1e-12)


def test_untruncated_pr_curve_matches_installed_sklearn():
    Y, S = _multilabel(300, 6, 1, ties=True)
    for k in range(Y.shape[1]):
        p, r, t = mo.precision_recall_curve(Y[:, k], S[:, k], truncate=False)
        ps, rs, ts = metrics.precision_recall_curve(Y[:, k], S[:, k])
        assert np.allclose(p, ps, rtol=0, atol=1e-15) and np.allclose(r, rs, rtol=0, atol=1e-15) and np.array_equal(t, ts)


def test_truncated_pr_curve_known_answer_of_sklearn_1_0():
    """The docstring example of sklearn 1.0.x precision_recall_curve (the release the reference pins):
    precision [0.667, 0.5, 1, 1], recall [1, 0.5, 0.5, 0], thresholds [0.35, 0.4, 0.8]."""
    y, s = np.array([0, 0, 1, 1]), np.array([0.1, 0.4, 0.35, 0.8])
    p, r, t = mo.precision_recall_curve(y, s, truncate=True)
    assert np.allclose(p, [2 / 3, 0.5, 1.0, 1.0]) and np.allclose(r, [1.0, 0.5, 0.5, 0.0]) and np.allclose(t, [0.35, 0.4, 0.8])
    p, r, t = mo.precision_recall_curve(y, s, truncate=False)          # releases >= 1.1 keep the whole curve
    assert np.allclose(p, [0.5, 2 / 3, 0.5, 1.0, 1.0]) and np.allclose(r, [1.0, 1.0, 0.5, 0.5, 0.0])


def test_degenerate_classes_follow_the_reference():
    """A class without positives: AP is NaN -> 0 and AUC raises -> 0, has_err set (loss_more.py:102-110)."""
    Y, S = _multilabel(200, 5, 2)
    Y[:, 3] = 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        text, per = mo.report(S, Y)
    assert text.count("Err(True)") == 1 and per["ap"][3] == 0.0 and per["auc"][3] == 0.0
    assert text.endswith("@ 200") and text.startswith("Mac-AP = ")
    with pytest.raises(ValueError):
        mo.roc_auc(Y[:, 3], S[:, 3])


def test_report_string_matches_sklearn_composition():
    """The whole report re-assembled from scikit-learn calls exactly as loss_more.py:92-130 does (installed release:
    un-truncated PR curve) equals the oracle's string with truncate=False."""
    Y, S = _multilabel(500, 20, 3)
    S = mo.zero_shot_scores(np.random.default_rng(4).standard_normal((500, 64)), np.random.default_rng(5).standard_normal((20, 64))) + 0.3 * Y
    ap_list, auc_list, ps, rs = [], [], [], []
    for k in range(Y.shape[1]):
        ap_list.append(metrics.average_precision_score(Y[:, k], S[:, k], average=None))
        auc_list.append(metrics.roc_auc_score(Y[:, k], S[:, k], average=None))
        p, r, _ = metrics.precision_recall_curve(Y[:, k], S[:, k])
        ps.append(p[len(p) // 2]); rs.append(r[len(p) // 2])
    want = (f"Mac-AP = {metrics.average_precision_score(Y, S, average='macro'):2.2f} "
            f"Mic-AP = {metrics.average_precision_score(Y, S, average='micro'):2.2f} "
            f"wAP = {metrics.average_precision_score(Y, S, average='weighted'):2.2f} "
            f"Err(False) mAP = {np.mean(ap_list) * 100.:2.2f} mAUC = {np.mean(auc_list) * 100.:2.2f} "
            f"mP = {np.mean(ps) * 100.:2.2f} mR = {np.mean(rs) * 100.:2.2f} @ 500")
    got, _ = mo.report(S, Y, truncate=False)
    assert got == want
