"""Host-side mirror of the reference loss-head API (vipant_b200/loss_head.py), CPU only.

Covers what does not need a GPU: construction / registry / checkpoint keys, the no-CPU-fallback contract, the
report-string arithmetic fed with the reference's golden ranks, and the row-sharded multi-process plumbing
(world_size 2, gloo) with an oracle-backed checker standing in for the CUDA kernel set.
"""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden
import vipant_b200 as vb
from vipant_b200 import _cabi, functional as F_
from oracle import infonce_oracle as io
from oracle.reference_loader import Cfg


def test_registry_and_ctor_contract():
    head = vb.build_loss_head(Cfg(name="CELossHead", scaling=True, scale_max=None))
    assert isinstance(head, vb.CELossHead) and isinstance(head, vb.LossHead)
    assert head.normalized is True and head.reduce is False
    assert isinstance(head.logit_scale, torch.nn.Parameter) and head.logit_scale.shape == ()
    assert head.logit_scale.item() == pytest.approx(math.log(1 / 0.07), rel=1e-6)
    assert list(head.state_dict().keys()) == ["logit_scale"]          # checkpoint key (cvap/model/cvap.py:112)
    assert head.scale_max == float("inf")
    assert vb.CELossHead(Cfg(scaling=True, scale_max=0)).scale_max == float("inf")      # `or inf` (:254)
    assert vb.CELossHead(Cfg(scaling=True, scale_max=100.0)).scale_max == 100.0
    fixed = vb.CELossHead(Cfg(scaling=False, scale_max=None))
    assert not isinstance(fixed.logit_scale, torch.nn.Parameter) and fixed.logit_scale.item() == 0.0
    assert list(fixed.state_dict().keys()) == []
    with pytest.raises(KeyError):
        vb.build_loss_head(Cfg(name="NoSuchHead"))


def test_copy_state_dict_roundtrip():
    head = vb.CELossHead(Cfg(scaling=True, scale_max=None))
    head.copy_state_dict({"logit_scale": torch.tensor(4.6052), "unrelated": torch.zeros(3)})
    assert head.logit_scale.item() == pytest.approx(4.6052)
    head2 = vb.CELossHead(Cfg(scaling=True, scale_max=None))
    head2.load_state_dict(head.state_dict())
    assert head2.logit_scale.item() == pytest.approx(4.6052)


def test_no_cpu_fallback():
    head = vb.CELossHead(Cfg(scaling=True, scale_max=None)).train()
    with pytest.raises(_cabi.VipantB200Error, match="no CPU fallback"):
        head(torch.randn(8, 64), torch.randn(8, 64))
    head.eval()
    with pytest.raises(_cabi.VipantB200Error, match="no CPU fallback"):
        head(torch.randn(8, 64), torch.randn(8, 64))
    with pytest.raises(_cabi.VipantB200Error):
        vb.sim_rank_topk(torch.randn(4, 64), torch.randn(5, 64), topk=1)


def test_composites_build_the_fused_head():
    val = vb.build_loss_head(Cfg(name="VALCELossHead", scaling=True, scale_max=None, va=False, lv=False, al=True))
    assert val.loss_head_va is None and val.loss_head_lv is None and isinstance(val.loss_head_al, vb.CELossHead)
    assert val.stats(nstep=1) == "al 0.000"
    va = vb.build_loss_head(Cfg(name="VACELossHead", scaling=True, scale_max=None, vp=True, ap=False, va=True, vv=False,
                                aa=False, vp_w=1., ap_w=1., va_w=1., vv_w=1., aa_w=1.))
    assert isinstance(va.loss_head_vp, vb.CELossHead) and va.loss_head_ap is None
    assert sorted(k for k in va.state_dict()) == ["loss_head_va.logit_scale", "loss_head_vp.logit_scale"]


def test_report_strings_from_golden_ranks(strings):
    g = load_golden("retrieval_1v5")
    r12 = torch.from_numpy(g["r12"].astype(np.int64))
    r21 = torch.from_numpy(g["r21"].astype(np.int64))
    ref_lines = strings["retrieval_1v5"].split("\nREFERENCE\n")[1]
    assert vb.LossHead._retrieval_eval_from_ranks(r12, r21) == ref_lines
    assert vb.LossHead.retrieval_metrics(r21.float(), msg="T->A") == ref_lines.split("\n")[1]


def test_install_into_reference():
    from oracle import reference_loader as rl
    if not rl.available():
        pytest.skip("reference tree not mounted (GPU box)")
    ref = rl.load_reference_loss_head()
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ref_copy_for_install", ref.__file__)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    vb.install_into_reference(mod)
    head = mod.build_loss_head(Cfg(name="CELossHead", scaling=True, scale_max=None))
    assert type(head) is vb.CELossHead
    comp = mod.build_loss_head(Cfg(name="VALCELossHead", scaling=True, scale_max=None, va=False, lv=False, al=True))
    assert type(comp) is vb.VALCELossHead


# ------------------------------------------------------------------- multi-process plumbing on gloo
class _CheckerKernels:
    """fp64 torch-CPU stand-in for functional._CudaKernels (same method contract), test-only: lets the
    all-gather / row-offset / scalar all-reduce logic of _sharded_forward/_sharded_backward run on gloo."""

    def normalize_pair(self, x1, x2, normalized, precision):
        x1, x2 = x1.double(), x2.double()
        if normalized:
            n1, n2 = torch.ones(len(x1), dtype=torch.float64), torch.ones(len(x2), dtype=torch.float64)
            a, t = x1, x2
        else:
            n1, n2 = x1.norm(dim=-1), x2.norm(dim=-1)
            a, t = x1 / n1[:, None], x2 / n2[:, None]
        return a, t, torch.stack([1 / n1, 1 / n2]), (a * t).sum(-1)

    def forward_sweep(self, a, t, a_all, t_all, row_offset, logit_scale, scale_max, precision, before_part2=None):
        if before_part2 is not None:
            before_part2()
        # single-pass regime: row sums are local, column sums are partial over the ranks -> all-reduced by the caller
        s, _ = io.effective_scale(float(logit_scale), scale_max)
        S = s * a @ t_all.T
        col_sum = torch.zeros((8, a_all.shape[0]), dtype=torch.float64)
        col_sum[0] = torch.exp(S - s).sum(0)
        return col_sum, (torch.logsumexp(S, 1), s)

    def forward_finish(self, b, B, D, row_offset, logit_scale, scale_max, dcos, precision, ws, col_sum):
        row_lse, s = ws
        _, flows = io.effective_scale(float(logit_scale), scale_max)
        col_lse = s + torch.log(col_sum.sum(0))[row_offset:row_offset + b]
        return torch.stack([row_lse, col_lse, s * dcos]), torch.tensor([s, float(flows)], dtype=torch.float64)

    def loss(self, stats_all):
        return ((stats_all[0] - stats_all[2]).mean() + (stats_all[1] - stats_all[2]).mean())

    def backward(self, x1, x2, a, t, a_all, t_all, inv, stats_all, scale, ws, row_offset, grad_out, normalized, precision):
        s, flows = float(scale[0]), float(scale[1])
        B, b = a_all.shape[0], a.shape[0]
        idx = torch.arange(b)
        S = s * a @ t_all.T
        G = torch.exp(S - stats_all[0][row_offset:row_offset + b, None]) + torch.exp(S - stats_all[1][None, :])
        G[idx, idx + row_offset] -= 2.0
        G *= float(grad_out) / B
        da = s * G @ t_all
        dls = (G * S).sum() * flows
        St = s * t @ a_all.T
        Gt = torch.exp(St - stats_all[1][row_offset:row_offset + b, None]) + torch.exp(St - stats_all[0][None, :])
        Gt[idx, idx + row_offset] -= 2.0
        Gt *= float(grad_out) / B
        dt = s * Gt @ a_all
        if not normalized:
            da = (da - a * (a * da).sum(-1, keepdim=True)) * inv[0][:, None]
            dt = (dt - t * (t * dt).sum(-1, keepdim=True)) * inv[1][:, None]
        return da, dt, dls


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, D, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x1n, x2n = io.make_pair(B, D, 0.3, 21)
        b = B // world
        x1 = torch.from_numpy(x1n[rank * b:(rank + 1) * b])
        x2 = torch.from_numpy(x2n[rank * b:(rank + 1) * b])
        ls = torch.tensor(math.log(1 / 0.07))
        kern = _CheckerKernels()
        loss, saved, extra = F_._sharded_forward(kern, x1, x2, ls, None, False, 0, dist.group.WORLD)
        dx1, dx2, dls = F_._sharded_backward(kern, saved, extra, torch.tensor(2.0), False, 0, dist.group.WORLD)
        out[rank] = (float(loss), dx1.numpy(), dx2.numpy(), float(dls))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_row_sharded_plumbing_gloo(world):
    B, D = 48, 32
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), B, D, out), nprocs=world, join=True)
    x1n, x2n = io.make_pair(B, D, 0.3, 21)
    ref = io.infonce_closed_form(x1n, x2n, float(np.float32(math.log(1 / 0.07))), grad_output=2.0)
    b = B // world
    for r in range(world):
        loss, dx1, dx2, dls = out[r]
        assert loss == pytest.approx(ref.loss, rel=1e-10)                      # same GLOBAL loss on every rank
        assert dls == pytest.approx(ref.dlogit_scale, rel=1e-9)                # all-reduced scalar
        np.testing.assert_allclose(dx1, ref.dx1[r * b:(r + 1) * b], atol=1e-12)
        np.testing.assert_allclose(dx2, ref.dx2[r * b:(r + 1) * b], atol=1e-12)


def test_gold_file_class_stats_match_reference(strings, tmp_path):
    """LossHead.report(gold_file=...) per-class line (reference :177-238) from nearest-neighbour indices."""
    from oracle.make_golden import gold_file_case
    a, t, ids, lines = gold_file_case()
    g = load_golden("retrieval_goldfile")
    path = tmp_path / "gold.json"
    path.write_text("\n".join(lines) + "\n")
    head = vb.CELossHead(Cfg(scaling=True, scale_max=None))
    head.ids = list(ids)
    by_class, by_sample = head._gold_cluster(str(path), len(ids))
    msg_12 = head._class_stats(torch.from_numpy(g["top1_12"].astype(np.int64)), by_class, by_sample, len(ids), "I->A")
    msg_21 = head._class_stats(torch.from_numpy(g["top1_21"].astype(np.int64)), by_class, by_sample, len(ids), "A->I")
    expected = strings["retrieval_nn_goldfile"].split("\n")[1]
    assert f"{msg_12} {msg_21}" == expected


# ---- host logic of the scoring reports with checker kernels standing in for the CUDA calls ---------------------------------
class _CheckerScoring:
    """numpy / fp64 stand-ins (the oracle's definitions) for the three CUDA entry points `report()` uses.  Installed with
    monkeypatch in the tests below ONLY: the package itself has no such path."""

    @staticmethod
    def l2_normalize(x, already_normalized=False):
        x = x.float()
        return x if already_normalized else x / x.norm(dim=-1, keepdim=True)

    @staticmethod
    def sim_rank_fused(q, k, gt_q=None, gt_k=None, top1_q=False, top1_k=False):
        from oracle import retrieval_oracle as ro
        S = q.double().numpy() @ k.double().numpy().T
        out = {}
        if gt_q is not None:
            out["ranks_q"] = torch.from_numpy(ro.rank_of(S, gt_q.numpy()))
        if gt_k is not None:
            out["ranks_k"] = torch.from_numpy(ro.rank_of(S.T, gt_k.numpy()))
        if top1_q:
            out["top1_q"] = (torch.from_numpy(S.argmax(1)), torch.from_numpy(S.max(1)).float())
        if top1_k:
            out["top1_k"] = (torch.from_numpy(S.argmax(0)), torch.from_numpy(S.max(0)).float())
        return out


@pytest.fixture
def checker_scoring(monkeypatch):
    monkeypatch.setattr(F_, "_require_cuda", lambda *a: None)
    monkeypatch.setattr(F_, "l2_normalize", _CheckerScoring.l2_normalize)
    monkeypatch.setattr(F_, "sim_rank_fused", _CheckerScoring.sim_rank_fused)


def test_report_host_logic_reproduces_the_reference_strings(checker_scoring, strings, tmp_path):
    """LossHead.infer / _stash / report (1-vs-5, N == M, N == M with a gold file, fallback shapes) fed by the checker: the
    strings must be the reference's (tests/golden/report_strings.json) -- this is the part of R1-R4 that is host Python."""
    from oracle.make_golden import gold_file_case, retrieval_inputs_1v5, retrieval_inputs_nn

    def run(a, t, batch, gold=None, names=None):
        head = vb.CELossHead(Cfg(scaling=True, scale_max=None)).eval()
        k = t.shape[0] // a.shape[0]
        for i in range(0, a.shape[0], batch):
            assert head(torch.from_numpy(a[i:i + batch]), torch.from_numpy(t[i * k:(i + batch) * k]), normalized=False,
                        names=None if names is None else names[i:i + batch]) is None
        rep = head.report(gold_file=gold)
        assert not hasattr(head, "x1s") and not hasattr(head, "_stash_normalized")
        return rep

    g = load_golden("retrieval_1v5_small")
    a, t = retrieval_inputs_1v5(n=150, seed=int(g["seed"]))
    assert run(a, t, 64) == strings["retrieval_1v5_small"]
    g = load_golden("retrieval_nn")
    a, t = retrieval_inputs_nn(seed=int(g["seed"]))
    assert run(a, t, 50) == strings["retrieval_nn"]
    a, t, ids, lines = gold_file_case()
    path = tmp_path / "gold.json"
    path.write_text("\n".join(lines) + "\n")
    assert run(a, t, 50, gold=str(path), names=ids) == strings["retrieval_nn_goldfile"]
    head = vb.CELossHead(Cfg(scaling=True, scale_max=None)).eval()
    torch.manual_seed(0)
    head(torch.randn(6, 16), torch.randn(9, 16))
    assert head.report() == strings["retrieval_fallback_6x9x16"]


def test_stash_keeps_views_compact():
    """A batch that is a view into something larger (the CLS token of a hidden state) is copied; a tensor that owns its
    storage is kept as it is (no extra kernel per batch)."""
    from vipant_b200.loss_head import _compact
    hidden = torch.randn(8, 5, 16)
    cls = hidden[:, 0, :]
    kept = _compact(cls)
    assert kept.is_contiguous() and kept.untyped_storage().nbytes() == 8 * 16 * 4 and torch.equal(kept, cls)
    own = torch.randn(8, 16)
    assert _compact(own).data_ptr() == own.data_ptr()


def test_bce_head_report_host_logic(monkeypatch):
    """BCELossHead.report: macro / weighted / mean fields and the Err flag from the per-class numbers (here supplied by the
    oracle) -- identical to the oracle's restatement of the reference's string, degenerate classes included."""
    from oracle import map_oracle as mo
    from vipant_b200 import loss_more as lm
    rng = np.random.default_rng(4)
    Y = (rng.random((400, 7)) < 0.2).astype(np.float32)
    Y[:, 2] = 0.0
    S = (Y + rng.standard_normal(Y.shape)).astype(np.float32)

    def fake(scores, labels, truncate_pr=True, micro=True):
        s, y = scores.numpy(), labels.numpy()
        _, parts = mo.report(s, y, truncate=truncate_pr)
        ap = np.array([mo.average_precision(y[:, k], s[:, k]) for k in range(y.shape[1])])
        auc = []
        for k in range(y.shape[1]):
            try:
                auc.append(mo.roc_auc(y[:, k], s[:, k]))
            except ValueError:
                auc.append(np.nan)
        return dict(ap=ap, auc=np.array(auc), p_mid=np.array(parts["p_mid"]), r_mid=np.array(parts["r_mid"]),
                    support=y.sum(0).astype(np.int64), flags=None, micro_ap=mo.average_precision_multilabel(y, s, "micro"))
    monkeypatch.setattr(lm, "multilabel_scores", fake)
    head = lm.BCELossHead(Cfg(embed_dim=16, width=16, layers=[8], bias=False, scaling=False), output_dim=7).eval()
    assert [type(m).__name__ for m in head.linear] == ["_LayerNormF32", "Linear", "_LayerNormF32", "Linear"]
    head.audios, head.x1s, head.x2s, head.ids = [], [], [], []
    got = head.report(x1s=torch.from_numpy(S), x2s=torch.from_numpy(Y))
    want, _ = mo.report(S, Y, truncate=True)
    assert got == want and "Err(True)" in got


def test_multi_pair_argument_checks():
    from vipant_b200 import functional as F
    assert not F.multi_pair_supported([torch.randn(4, 512)])                       # CPU tensors: no fused step, no fallback
    with pytest.raises(_cabi.VipantB200Error):
        vb.infonce_multi_loss([torch.randn(4, 512), torch.randn(4, 512)], [(0, 1)], [torch.tensor(1.0)])
    with pytest.raises(ValueError):
        F.check_gt_range(torch.tensor([[0, 7]]), 7)
    F.check_gt_range(torch.tensor([[0, 6]]), 7)
    with pytest.raises(ValueError):
        F._gt_matrix(torch.zeros(3, 9, dtype=torch.long), 3, 10, torch.device("cpu"), "gt")


def test_fused_post_encoder_mirrors_reference_parameters_and_refuses_cpu():
    """FusedPostEncoder keeps the reference's checkpoint keys (ViTPostEncoder: `ln`, `proj`, cvap/module/val.py:270-273)
    and, like every product path here, has no CPU fallback."""
    import torch
    from vipant_b200._cabi import VipantB200Error
    from vipant_b200.encoder_tail import FusedPostEncoder, encoder_tail
    m = FusedPostEncoder(None, width=768, embed_dim=512)
    assert sorted(m.state_dict()) == ["ln.bias", "ln.weight", "proj"]
    assert tuple(m.proj.shape) == (768, 512) and m.ln.normalized_shape == (768,)
    with pytest.raises(VipantB200Error):
        m(torch.randn(4, 3, 768))
    with pytest.raises(VipantB200Error):
        encoder_tail(torch.randn(4, 768), m.ln.weight, m.ln.bias, m.proj)


def test_encoder_tail_backward_formulas_match_the_oracle():
    """tail_backward (the torch side of the fused tail's autograd: two GEMMs + LayerNorm backward from kept statistics) against
    the float64 closed forms of oracle/encoder_tail_oracle.py, on CPU tensors (the function is device-agnostic; the kernels'
    forward outputs are stood in for by the oracle's)."""
    import numpy as np
    import torch
    from oracle import encoder_tail_oracle as eo
    from vipant_b200.encoder_tail import tail_backward
    inp = eo.golden_inputs(seed=99, rows=48, tokens=1, width=256, embed=256)
    x = inp["hidden"][:, 0, :]
    x64 = x.astype(np.float64)
    mean = x64.mean(-1)
    rstd = 1.0 / np.sqrt(((x64 - mean[:, None]) ** 2).mean(-1) + eo.EPS)
    ln = eo.layer_norm(x, inp["gamma"], inp["beta"])
    y, unit = eo.encoder_tail(x, inp["gamma"], inp["beta"], inp["proj"])
    w = inp["w"].astype(np.float64)
    nrm = np.linalg.norm(y, axis=-1, keepdims=True)
    dy = (w - unit * (w * unit).sum(-1, keepdims=True)) / nrm                       # d sum(w * y/||y||) / dy
    t = lambda a: torch.from_numpy(np.asarray(a, np.float32))                       # noqa: E731
    got = tail_backward(t(x), t(inp["gamma"]), t(inp["proj"]), t(ln), t(mean), t(rstd), t(dy))
    want = eo.encoder_tail_grads(x, inp["gamma"], inp["beta"], inp["proj"], inp["w"])
    for g, r, name in zip(got, want, ("x", "ln.weight", "ln.bias", "proj")):
        rel = np.linalg.norm(g.numpy() - r) / np.linalg.norm(r)
        assert rel < 2e-5, (name, rel)
