"""Property-based checks (hypothesis) of the oracle restatements and the host-side formats: random shapes, ties, ragged items."""
import os

import numpy as np
import pytest
import torch

hyp = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st  # noqa: E402

from oracle import infonce_oracle as io  # noqa: E402
from oracle import map_oracle as mo  # noqa: E402
from oracle import retrieval_oracle as ro  # noqa: E402
from vipant_b200 import embed_cache as ec  # noqa: E402

FAST = settings(max_examples=25, deadline=None)


@FAST
@given(st.integers(2, 24), st.integers(1, 40), st.integers(0, 2 ** 31 - 1))
def test_rank_oracle_equals_position_in_descending_sort(n, m, seed):
    """rank_of == position of the column in torch's descending argsort on tie-free rows (what loss_head.py:115-117 computes)."""
    rng = np.random.default_rng(seed)
    S = rng.permutation(n * m).reshape(n, m).astype(np.float32)          # all distinct: argsort is unambiguous
    gt = rng.integers(0, m, size=n)
    order = torch.from_numpy(S).argsort(descending=True).numpy()
    want = np.array([np.where(order[i] == gt[i])[0][0] for i in range(n)])
    assert np.array_equal(ro.rank_of(S, gt)[:, 0], want)
    idx, val = ro.topk(S, min(5, m))
    assert np.array_equal(idx, order[:, :min(5, m)])


@FAST
@given(st.integers(2, 40), st.integers(4, 48), st.floats(0.5, 120.0), st.integers(0, 2 ** 31 - 1))
def test_closed_form_matches_torch_autograd(B, D, s, seed):
    """The fp64 closed form (loss and all three gradients) against autograd of the reference's formula, any temperature."""
    rng = np.random.default_rng(seed)
    x1n, x2n = rng.standard_normal((B, D)), rng.standard_normal((B, D))
    ref = io.infonce_closed_form(x1n, x2n, float(np.log(s)), None, False, 1.7)
    x1 = torch.tensor(x1n, requires_grad=True)
    x2 = torch.tensor(x2n, requires_grad=True)
    ls = torch.tensor(float(np.log(s)), dtype=torch.float64, requires_grad=True)
    a, t = x1 / x1.norm(dim=-1, keepdim=True), x2 / x2.norm(dim=-1, keepdim=True)
    lab = torch.arange(B)
    loss = torch.nn.functional.cross_entropy(ls.exp() * a @ t.t(), lab) + torch.nn.functional.cross_entropy(ls.exp() * t @ a.t(), lab)
    (1.7 * loss).backward()
    assert abs(ref.loss - loss.item()) <= 1e-9 * max(1.0, abs(loss.item()))
    assert np.allclose(ref.dx1, x1.grad.numpy(), rtol=1e-7, atol=1e-10) and np.allclose(ref.dx2, x2.grad.numpy(), rtol=1e-7, atol=1e-10)
    assert abs(ref.dlogit_scale - ls.grad.item()) <= 1e-7 * max(1.0, abs(ls.grad.item()))


@FAST
@given(st.integers(5, 200), st.integers(0, 2 ** 31 - 1), st.booleans())
def test_ap_auc_oracle_matches_sklearn(n, seed, ties):
    metrics = pytest.importorskip("sklearn.metrics")
    rng = np.random.default_rng(seed)
    y = (rng.random(n) < 0.3).astype(int)
    if y.sum() in (0, n):
        y[0], y[-1] = 1, 0
    s = rng.standard_normal(n)
    if ties:
        s = np.round(s)
    assert mo.average_precision(y, s) == pytest.approx(metrics.average_precision_score(y, s), rel=1e-12)
    assert mo.roc_auc(y, s) == pytest.approx(metrics.roc_auc_score(y, s), rel=1e-12)


@FAST
@given(st.lists(st.tuples(st.integers(1, 6), st.booleans()), min_size=0, max_size=12), st.integers(1, 40), st.booleans(),
       st.integers(0, 2 ** 31 - 1))
def test_embedding_shard_round_trip(items, dim, bf16, seed):
    import tempfile
    rng = np.random.default_rng(seed)
    data = []
    for i, (k, vector) in enumerate(items):
        arr = rng.standard_normal(dim).astype(np.float32) if vector else rng.standard_normal((k, dim)).astype(np.float32)
        data.append((f"item/{i}", arr))
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "s.vpae")
        ec.pack_items(data, path, ec.DTYPE_BF16 if bf16 else ec.DTYPE_F32)
        shard = ec.EmbeddingShard(path)
        assert shard.names == [n for n, _ in data]
        for name, arr in data:
            want = torch.from_numpy(arr).to(torch.bfloat16).float().numpy() if bf16 else arr
            got = shard[name]
            assert got.shape == arr.shape and np.array_equal(got, want)
        if data:
            order = [data[i][0] for i in rng.permutation(len(data))]
            batch, counts = shard.gather(order)
            assert batch.shape[0] == counts.sum() == sum(np.atleast_2d(dict(data)[n]).shape[0] for n in order)
        del shard
