"""GPU parity of the AudioSet multi-label scoring (SURVEY.md 8f row 3 = Z2) against oracle/map_oracle.py, the restatement of the
scikit-learn algorithms the reference calls (loss_more.py:92-123), pinned to scikit-learn in tests/test_oracle_map.py.
Bars: fp64 sums in a different (fixed) order -> 1e-9 relative; the report string must be identical."""
import numpy as np
import pytest
import torch

from oracle import map_oracle as mo
from oracle.reference_loader import Cfg

pytestmark = pytest.mark.gpu


def _case(N, C, density, seed, levels=None, signal=1.0):
    rng = np.random.default_rng(seed)
    Y = (rng.random((N, C)) < density).astype(np.float32)
    S = (signal * Y + rng.standard_normal((N, C))).astype(np.float32)
    if levels:
        S = (np.round(S * levels) / levels).astype(np.float32)      # many exact ties: thresholds group several samples
    return S, Y


def _oracle(S, Y, truncate):
    C = S.shape[1]
    ap, auc, pm, rm = np.full(C, np.nan), np.full(C, np.nan), np.zeros(C), np.zeros(C)
    for k in range(C):
        ap[k] = mo.average_precision(Y[:, k], S[:, k])
        try:
            auc[k] = mo.roc_auc(Y[:, k], S[:, k])
        except ValueError:
            pass
        p, r, _ = mo.precision_recall_curve(Y[:, k], S[:, k], truncate=truncate)
        pm[k], rm[k] = p[len(p) // 2], r[len(p) // 2]
    return ap, auc, pm, rm


def _close(a, b, tol=1e-9):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    ok = ~np.isnan(b)
    assert np.all(np.abs(a[ok] - b[ok]) <= tol * np.maximum(1.0, np.abs(b[ok]))), float(np.abs(a[ok] - b[ok]).max())


@pytest.mark.parametrize("N,C,density,levels", [(3000, 40, 0.03, None), (1500, 17, 0.3, 4), (1000, 5, 0.5, 1), (33, 3, 0.4, None),
                                                (1025, 9, 0.02, 8), (4096, 12, 0.1, None)])
@pytest.mark.parametrize("truncate", [True, False])
def test_per_class_and_micro_against_oracle(N, C, density, levels, truncate):
    from vipant_b200.loss_more import multilabel_scores
    S, Y = _case(N, C, density, 100 + N + C, levels)
    Y[0, :] = 1.0                                        # every class has a positive ...
    Y[1, :] = 0.0                                        # ... and a negative
    m = multilabel_scores(torch.from_numpy(S).cuda(), torch.from_numpy(Y).cuda(), truncate_pr=truncate)
    ap, auc, pm, rm = _oracle(S, Y, truncate)
    _close(m["ap"], ap); _close(m["auc"], auc); _close(m["p_mid"], pm); _close(m["r_mid"], rm)
    _close([m["micro_ap"]], [mo.average_precision_multilabel(Y, S, "micro")])
    assert np.array_equal(m["support"], Y.sum(0).astype(np.int64)) and not m["flags"].any()
    # uint8 / bool labels give the same numbers
    m8 = multilabel_scores(torch.from_numpy(S).cuda(), torch.from_numpy(Y).cuda().bool(), truncate_pr=truncate)
    assert np.array_equal(m8["ap"], m["ap"]) and m8["micro_ap"] == m["micro_ap"]


def test_degenerate_classes_and_report_string():
    """Classes without a positive (AP undefined -> 0 + Err) or without a negative (AUC undefined), constant scores; the report
    string of BCELossHead.report equals the oracle's restatement of the reference's."""
    import vipant_b200  # noqa: F401
    from vipant_b200.loss_more import BCELossHead, multilabel_scores
    S, Y = _case(800, 6, 0.2, 7)
    Y[:, 1] = 0.0                  # no positive
    Y[:, 2] = 1.0                  # no negative
    S[:, 3] = 0.25                 # one threshold only
    m = multilabel_scores(torch.from_numpy(S).cuda(), torch.from_numpy(Y).cuda())
    ap, auc, pm, rm = _oracle(S, Y, True)
    _close(m["ap"], ap); _close(m["auc"], auc); _close(m["p_mid"], pm); _close(m["r_mid"], rm)
    assert m["flags"].tolist() == [0, 1, 2, 0, 0, 0] and np.isnan(m["ap"][1]) and np.isnan(m["auc"][1]) and np.isnan(m["auc"][2])
    want, _ = mo.report(S, Y, truncate=True)
    head = BCELossHead(Cfg(embed_dim=16, width=16, layers=[], bias=True, scaling=True), output_dim=6).cuda().eval()
    head.audios, head.x1s, head.x2s, head.ids = [], [], [], []
    got = head.report(x1s=torch.from_numpy(S).cuda(), x2s=torch.from_numpy(Y).cuda())
    assert got == want, (got, want)
    assert not hasattr(head, "x1s")


def test_audioset_shape_zero_shot():
    """AudioSet evaluation shape: 20371 clips x 527 label prompts, ~2 labels per clip.  BCELossHead.zero_shot end to end
    (normalise, similarity, scoring) against the oracle on the same embeddings."""
    import vipant_b200  # noqa: F401
    from vipant_b200.loss_more import BCELossHead
    N, C, D = 20371, 527, 512
    rng = np.random.default_rng(1213)
    text = rng.standard_normal((C, D)).astype(np.float32)
    Y = np.zeros((N, C), np.float32)
    for _ in range(2):
        Y[np.arange(N), rng.integers(0, C, N)] = 1.0
    audios = (Y @ text * 0.08 + rng.standard_normal((N, D))).astype(np.float32)
    head = BCELossHead(Cfg(embed_dim=D, width=D, layers=[], bias=True, scaling=True), output_dim=C).cuda().eval()
    with torch.no_grad():
        for i in range(0, N, 4096):
            loss = head(torch.from_numpy(audios[i:i + 4096]).cuda(), torch.from_numpy(Y[i:i + 4096]).cuda(), names=None)
            assert loss.dim() == 0
        got = head.report(gold_file=None, text=torch.from_numpy(text).cuda())
    S = mo.zero_shot_scores(audios, text)
    want, parts = mo.report(S, Y, truncate=True)
    # the similarities are fp32 sums in another order than numpy's: the metrics agree to ~1e-5, and for this seed the
    # two-decimal report fields are identical
    fields = lambda s: [float(x) for x in __import__("re").findall(r"= (-?\d+\.\d+|nan)", s)]
    assert len(fields(got)) == len(fields(want)) == 7
    assert np.allclose(fields(got), fields(want), atol=0.011), (got, want)
    assert got.endswith(f"@ {N}") and got.startswith("Mac-AP")


def test_too_many_samples_is_an_error():
    from vipant_b200 import _cabi
    from vipant_b200.loss_more import multilabel_scores
    with pytest.raises(_cabi.VipantB200Error):
        multilabel_scores(torch.zeros(32769, 2, device="cuda"), torch.zeros(32769, 2, device="cuda"))
