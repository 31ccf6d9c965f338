"""The CPU oracle against the vectors the UNMODIFIED reference produced (tests/golden, oracle/make_golden.py).

CPU only.  These tests pin the restatement (oracle/*.py) that the GPU parity tests then use as the checker.
Tolerances: the golden vectors are fp32 reference outputs, the closed form runs in fp64 -> 2e-6 relative on the
loss, 2e-4 relative (Frobenius) on gradients (fp32 rounding of a 2048-term softmax/matmul chain).
"""
import math
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from oracle import infonce_oracle as io
from oracle import retrieval_oracle as ro
from oracle.make_golden import (INFONCE_CASES, checksum, infonce_inputs, retrieval_inputs_1v5, retrieval_inputs_nn,
                                zero_shot_inputs)


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("name", sorted(INFONCE_CASES))
def test_infonce_closed_form_matches_reference(name):
    case = INFONCE_CASES[name]
    g = load_golden("infonce_" + name)
    x1, x2 = infonce_inputs(case)
    assert checksum(x1, x2) == pytest.approx(float(g["input_checksum"]), rel=1e-12), "input generator drifted"
    res = io.infonce_closed_form(x1, x2, case["logit_scale"], case["scale_max"], case["normalized"], case["grad_output"])
    assert res.loss == pytest.approx(float(g["loss"]), rel=2e-6)
    assert res.dlogit_scale == pytest.approx(float(g["dlogit_scale"]), rel=2e-4, abs=1e-7 * case["grad_output"])
    assert np.linalg.norm(res.dx1) == pytest.approx(float(g["dx1_norm"]), rel=2e-4)
    assert np.linalg.norm(res.dx2) == pytest.approx(float(g["dx2_norm"]), rel=2e-4)
    if "dx1" in g.files:
        assert rel(res.dx1, g["dx1"]) < 2e-4 and rel(res.dx2, g["dx2"]) < 2e-4
        assert np.array_equal(x1, g["x1"]) and np.array_equal(x2, g["x2"])
    else:
        rows = g["rows"]
        assert rel(res.dx1[rows], g["dx1_rows"]) < 2e-4 and rel(res.dx2[rows], g["dx2_rows"]) < 2e-4


def test_anchor_from_survey():
    """torch.manual_seed(0), randn(64,512) x2 -> loss 8.825647354125977, dls 0.9049414."""
    import torch
    torch.manual_seed(0)
    x1, x2 = torch.randn(64, 512), torch.randn(64, 512)
    res = io.infonce_closed_form(x1.numpy(), x2.numpy())
    assert res.loss == pytest.approx(8.825647354125977, rel=1e-6)
    assert res.dlogit_scale == pytest.approx(0.9049414, rel=1e-5)
    assert np.linalg.norm(res.dx1) == pytest.approx(0.157243, rel=1e-4)


def test_clamp_stops_scale_gradient():
    case = INFONCE_CASES["b200_prenorm_clamped"]
    x1, x2 = infonce_inputs(case)
    res = io.infonce_closed_form(x1, x2, case["logit_scale"], case["scale_max"], True)
    assert res.scale == 100.0 and res.dlogit_scale == 0.0
    assert float(load_golden("infonce_b200_prenorm_clamped")["dlogit_scale"]) == 0.0
    # scale_max = 0 / None -> no clamp (`cfg.scale_max or float("inf")`, loss_head.py:254)
    assert io.effective_scale(math.log(200.0), 0) == (pytest.approx(200.0), True)
    assert io.effective_scale(math.log(200.0), None) == (pytest.approx(200.0), True)
    assert io.effective_scale(math.log(50.0), 100.0) == (pytest.approx(50.0), True)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_row_sharded_decomposition_equals_global_batch(world):
    x1, x2 = io.make_pair(256, 128, 0.3, 5)
    full = io.infonce_closed_form(x1, x2, grad_output=3.0)
    sh = io.infonce_row_sharded(x1, x2, world, grad_output=3.0)
    assert sh.loss == pytest.approx(full.loss, rel=1e-13)
    assert sh.dlogit_scale == pytest.approx(full.dlogit_scale, rel=1e-11)
    np.testing.assert_allclose(sh.dx1, full.dx1, rtol=0, atol=1e-15)
    np.testing.assert_allclose(sh.dx2, full.dx2, rtol=0, atol=1e-15)
    np.testing.assert_allclose(sh.row_lse, full.row_lse, rtol=1e-14)
    np.testing.assert_allclose(sh.col_lse, full.col_lse, rtol=1e-14)


def test_port_torch_matches_closed_form():
    import torch
    x1n, x2n = io.make_pair(96, 64, 0.2, 9)
    x1 = torch.from_numpy(x1n).double().requires_grad_(True)
    x2 = torch.from_numpy(x2n).double().requires_grad_(True)
    ls = torch.tensor(2.0, dtype=torch.float64, requires_grad=True)
    loss = io.infonce_port_torch(x1, x2, ls)
    loss.backward()
    res = io.infonce_closed_form(x1n, x2n, 2.0)
    assert loss.item() == pytest.approx(res.loss, rel=1e-12)
    assert rel(x1.grad.numpy(), res.dx1) < 1e-10 and rel(x2.grad.numpy(), res.dx2) < 1e-10
    assert ls.grad.item() == pytest.approx(res.dlogit_scale, rel=1e-10)


# ---------------------------------------------------------------- retrieval / zero-shot
@pytest.mark.parametrize("tag,n", [("retrieval_1v5", 975), ("retrieval_1v5_small", 150)])
def test_retrieval_1v5_matches_reference(tag, n, strings):
    g = load_golden(tag)
    a, t = retrieval_inputs_1v5(n=n, seed=int(g["seed"]))
    assert checksum(a, t) == pytest.approx(float(g["input_checksum"]), rel=1e-12)
    rep, r12, r21 = ro.report(ro.normalize(a), ro.normalize(t))
    assert rep == strings[tag]
    assert np.array_equal(r12, g["r12"]) and np.array_equal(r21, g["r21"])
    S = ro.similarity(ro.normalize(a), ro.normalize(t))
    idx, _ = ro.topk(S, 10)
    ok = g["top10_ok"]
    assert np.array_equal(idx[ok], g["top10"][ok])


def test_retrieval_nn_matches_reference(strings):
    g = load_golden("retrieval_nn")
    a, t = retrieval_inputs_nn(seed=int(g["seed"]))
    rep, r12, r21 = ro.report(ro.normalize(a), ro.normalize(t))
    assert rep == strings["retrieval_nn"]
    assert np.array_equal(r12, g["r12"]) and np.array_equal(r21, g["r21"])


def test_retrieval_fallback_string(strings):
    rep, _, _ = ro.report(np.zeros((6, 16), np.float32), np.zeros((9, 16), np.float32))
    assert rep == strings["retrieval_fallback_6x9x16"]


@pytest.mark.parametrize("c,tag", [(50, "zs50"), (200, "zs200")])
def test_zero_shot_matches_reference(c, tag, strings):
    g = load_golden("zero_shot_" + tag)
    audios, text, labels = zero_shot_inputs(c=c, seed=int(g["seed"]))
    label_map = {i: i // 4 for i in range(200)} if c == 200 else None
    rep, pred = ro.zero_shot_report(audios, text, labels, label_map)
    assert rep == strings["zero_shot_" + tag]
    assert np.array_equal(pred, g["pred"])


def test_rank_and_topk_tie_semantics():
    S = np.array([[1.0, 3.0, 3.0, 2.0, 3.0]], np.float32)
    # strictly-greater count (what argsort+where gives on tie-free rows)
    assert ro.rank_of(S, np.array([[3]]))[0, 0] == 3
    idx, val = ro.topk(S, 3)
    assert idx.tolist() == [[1, 2, 4]] and val.tolist() == [[3.0, 3.0, 3.0]]


def test_bench_parity_fixture_generator_matches_closed_form():
    """oracle/make_bench_parity.py (the fp64 ground truth bench.py's `parity` object is checked against) evaluates the
    logits blockwise; on a size the dense closed form handles it must give the same numbers, and the committed
    B = 32768 fixture must be the one the generator's sampling scheme describes."""
    import torch
    from oracle import make_bench_parity as mb
    x1n, x2n = io.make_pair(640, 512, 0.3, 1213)
    ref = io.infonce_closed_form(x1n.astype(np.float64), x2n.astype(np.float64), float(np.float32(np.log(1 / 0.07))))
    loss, dls, dx1, dx2 = mb.blockwise_truth(torch.from_numpy(x1n), torch.from_numpy(x2n), rb=256)
    # (the generator takes s = expf(fp32 parameter) as the kernels do, the closed form exp() in double: 1e-8 apart)
    assert abs(loss - ref.loss) < 1e-6 * abs(ref.loss)
    assert abs(dls - ref.dlogit_scale) < 1e-6 * abs(ref.dlogit_scale)
    assert np.linalg.norm(dx1.numpy() - ref.dx1) < 1e-6 * np.linalg.norm(ref.dx1)
    assert np.linalg.norm(dx2.numpy() - ref.dx2) < 1e-6 * np.linalg.norm(ref.dx2)
    fx = np.load(os.path.join(GOLDEN, "bench_parity_b32768.npz"))
    assert int(fx["B"]) == 32768 and int(fx["block"]) == mb.BLOCK
    assert np.array_equal(fx["rows"], mb.sample_rows(32768)) and fx["dx1_rows"].shape == (8 * mb.SAMPLES, 512)
    assert fx["dx1_block_norm"].shape == (8,) and abs(float(fx["loss"]) - 9.96007) < 1e-4


# ---------------------------------------------------------------- encoder tail
ENCODER_TAIL_CASES = {"vit": dict(seed=4101, rows=72, tokens=3, width=768, embed=512),
                      "gpt": dict(seed=4102, rows=40, tokens=6, width=512, embed=512),
                      "vit256": dict(seed=4103, rows=130, tokens=2, width=1024, embed=256)}


@pytest.mark.parametrize("name", sorted(ENCODER_TAIL_CASES))
def test_encoder_tail_oracle_matches_reference_post_encoders(name):
    """The numpy restatement against the outputs of the reference's own ViTPostEncoder / GPTPostEncoder (+ the heads'
    normalisation), fp32 on CPU: 2e-5 of a row's norm (fp32 accumulation over width <= 1024 against float64)."""
    from oracle import encoder_tail_oracle as eo
    fx = load_golden("encoder_tail")
    inp = eo.golden_inputs(**ENCODER_TAIL_CASES[name])
    sums = [float(np.asarray(inp[k], np.float64).sum()) for k in ("hidden", "gamma", "beta", "proj")]
    assert np.allclose(sums, fx[f"{name}_checksum"], rtol=1e-12, atol=0)              # same inputs as the generator's
    y, unit = eo.encoder_tail(inp["hidden"], inp["gamma"], inp["beta"], inp["proj"], mask=inp["eot"] if name == "gpt" else None)
    norm = fx[f"{name}_norm"].astype(np.float64)
    assert (np.linalg.norm(y - fx[f"{name}_y"], axis=-1) / norm).max() < 2e-5
    assert np.abs(np.linalg.norm(y, axis=-1) / norm - 1).max() < 1e-5
    assert np.abs(unit[::8] - fx[f"{name}_unit_rows"]).max() < 2e-6
    if name == "vit":
        from oracle.make_golden_encoder_tail import GRAD_ROWS
        dx, dgamma, dbeta, dproj = eo.encoder_tail_grads(inp["hidden"][:, 0, :], inp["gamma"], inp["beta"], inp["proj"], inp["w"])
        assert float(fx["vit_dx_other_tokens_absmax"]) == 0.0                          # only the CLS token carries gradient
        assert abs(np.linalg.norm(dx) / float(fx["vit_dx_norm"]) - 1) < 1e-5
        assert np.linalg.norm(dx[GRAD_ROWS] - fx["vit_dx_rows"]) < 2e-5 * np.linalg.norm(fx["vit_dx_rows"])
        assert np.linalg.norm(dgamma - fx["vit_dgamma"]) < 2e-5 * np.linalg.norm(fx["vit_dgamma"])
        assert np.linalg.norm(dbeta - fx["vit_dbeta"]) < 2e-5 * np.linalg.norm(fx["vit_dbeta"])
        assert abs(np.linalg.norm(dproj) / float(fx["vit_dproj_norm"]) - 1) < 1e-5
        assert np.linalg.norm(dproj[GRAD_ROWS] - fx["vit_dproj_rows"]) < 2e-5 * np.linalg.norm(fx["vit_dproj_rows"])


def test_encoder_tail_golden_generator_reproduces_fixture():
    """With the reference mounted (build container), running its post-encoders again gives the committed fixture."""
    from oracle import reference_loader as rl
    if not rl.available():
        pytest.skip("reference not mounted")
    from oracle import make_golden_encoder_tail as mg
    assert mg.CASES == ENCODER_TAIL_CASES
    res = mg.build()
    fx = load_golden("encoder_tail")
    assert sorted(res) == sorted(fx.files)
    for k in fx.files:
        np.testing.assert_allclose(res[k], fx[k], rtol=1e-6, atol=1e-7)
