"""GPU parity of the InfoNCE path (normalise -> forward statistics -> loss -> backward) through the C-ABI.

Checker: the fp64 closed form of oracle/infonce_oracle.py (pinned to the reference in test_oracle.py) and the
committed reference outputs in tests/golden/.  Tolerances are BASELINE.json's:
  fp32 mode  loss 1e-4 rel, gradients 1e-3 rel (Frobenius);  bf16 mode  loss 1e-3 rel, gradients 1e-2 rel.
"""
import ctypes
import math

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import infonce_oracle as io
from oracle.make_golden import INFONCE_CASES, infonce_inputs
from oracle.reference_loader import Cfg

pytestmark = pytest.mark.gpu

TOL = {"fp32": dict(loss=1e-4, grad=1e-3), "bf16": dict(loss=1e-3, grad=1e-2)}


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def run(x1n, x2n, precision, logit_scale=math.log(1 / 0.07), scale_max=None, normalized=False, grad_output=1.0,
        dtype=torch.float32):
    import vipant_b200 as vb
    x1 = torch.from_numpy(x1n).cuda().to(dtype).requires_grad_(True)
    x2 = torch.from_numpy(x2n).cuda().to(dtype).requires_grad_(True)
    ls = torch.tensor(logit_scale, device="cuda", dtype=torch.float32, requires_grad=True)
    loss = vb.infonce_loss(x1, x2, ls, scale_max=scale_max, normalized=normalized, precision=precision)
    (loss * grad_output).backward()
    return loss.item(), x1.grad.float().cpu().numpy(), x2.grad.float().cpu().numpy(), ls.grad.item()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", sorted(INFONCE_CASES))
def test_golden_cases(name, precision):
    """Same seeded inputs as the reference run that produced tests/golden/infonce_<name>.npz."""
    case = INFONCE_CASES[name]
    g = load_golden("infonce_" + name)
    x1n, x2n = infonce_inputs(case)
    tol = TOL[precision]
    loss, dx1, dx2, dls = run(x1n, x2n, precision, case["logit_scale"], case["scale_max"], case["normalized"],
                              case["grad_output"])
    ref = io.infonce_closed_form(x1n, x2n, case["logit_scale"], case["scale_max"], case["normalized"], case["grad_output"])
    # against the reference's own outputs
    assert abs(loss - float(g["loss"])) <= tol["loss"] * abs(float(g["loss"]))
    if "dx1" in g.files:
        assert rel(dx1, g["dx1"]) <= tol["grad"] and rel(dx2, g["dx2"]) <= tol["grad"]
    else:
        rows = g["rows"]
        assert rel(dx1[rows], g["dx1_rows"]) <= tol["grad"] and rel(dx2[rows], g["dx2_rows"]) <= tol["grad"]
    # against the fp64 oracle, all rows
    assert abs(loss - ref.loss) <= tol["loss"] * abs(ref.loss)
    assert rel(dx1, ref.dx1) <= tol["grad"] and rel(dx2, ref.dx2) <= tol["grad"]
    scale_ref = max(abs(ref.dlogit_scale), 1e-3 * case["grad_output"])
    assert abs(dls - ref.dlogit_scale) <= tol["grad"] * scale_ref
    if name == "b200_prenorm_clamped":
        assert dls == 0.0            # saturated clamp passes no gradient (torch.clamp semantics)


@pytest.mark.parametrize("precision,B,D", [
    ("fp32", 1, 64), ("fp32", 7, 36), ("fp32", 129, 100), ("fp32", 300, 768),
    ("bf16", 1, 64), ("bf16", 7, 128), ("bf16", 129, 192), ("bf16", 257, 256), ("bf16", 300, 384), ("bf16", 1000, 512),
    ("bf16", 128, 512), ("bf16", 4096, 512),
])
def test_ragged_shapes(precision, B, D):
    x1n, x2n = io.make_pair(B, D, 0.3, 100 + B)
    ref = io.infonce_closed_form(x1n, x2n, grad_output=0.5)
    loss, dx1, dx2, dls = run(x1n, x2n, precision, grad_output=0.5)
    tol = TOL[precision]
    # tiny batches: the loss is a small difference of O(10) logsumexps, so the relative bar gets an absolute floor
    assert abs(loss - ref.loss) <= tol["loss"] * max(abs(ref.loss), 1.0)
    if B == 1:      # a single pair has zero loss and zero gradient: absolute check
        assert np.abs(dx1).max() <= 1e-5 and np.abs(dx2).max() <= 1e-5 and abs(dls) <= 1e-5
        return
    assert rel(dx1, ref.dx1) <= tol["grad"] and rel(dx2, ref.dx2) <= tol["grad"]
    assert abs(dls - ref.dlogit_scale) <= tol["grad"] * max(abs(ref.dlogit_scale), 1e-3)


@pytest.mark.parametrize("s", [5.0, 42.9, 43.2, 60.0, 120.0])
def test_forward_regimes(s):
    """The tensor-core forward picks, ON THE DEVICE, the single-pass kernel (s*log2e <= 62, i.e. s <= 42.97) or the
    exact two-sweep kernel; both sides of the switch and a temperature far above it must meet the same bars."""
    x1n, x2n = io.make_pair(777, 512, 0.05, 17)         # weakly aligned pairs: the loss stays O(1) at every temperature
    ref = io.infonce_closed_form(x1n, x2n, math.log(s))
    loss, dx1, dx2, dls = run(x1n, x2n, "bf16", logit_scale=math.log(s))
    assert abs(loss - ref.loss) <= 1e-3 * max(abs(ref.loss), 0.1)
    assert rel(dx1, ref.dx1) <= 1e-2 and rel(dx2, ref.dx2) <= 1e-2
    assert abs(dls - ref.dlogit_scale) <= 1e-2 * abs(ref.dlogit_scale)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_half_inputs(dtype):
    """Encoders under autocast hand over 16-bit features; gradients come back in the same dtype."""
    x1n, x2n = io.make_pair(384, 512, 0.3, 5)
    x1h = torch.from_numpy(x1n).to(dtype).float().numpy()       # the values the kernel actually sees
    x2h = torch.from_numpy(x2n).to(dtype).float().numpy()
    ref = io.infonce_closed_form(x1h, x2h)
    loss, dx1, dx2, dls = run(x1n, x2n, "bf16", dtype=dtype)
    assert abs(loss - ref.loss) <= 1e-3 * abs(ref.loss)
    assert rel(dx1, ref.dx1) <= 1.5e-2 and rel(dx2, ref.dx2) <= 1.5e-2      # + one rounding of the output to 16 bit


def test_gradient_is_linear_in_grad_output_and_tangent():
    x1n, x2n = io.make_pair(1024, 512, 0.3, 77)
    _, dx1a, dx2a, dlsa = run(x1n, x2n, "bf16", grad_output=1.0)
    _, dx1b, dx2b, dlsb = run(x1n, x2n, "bf16", grad_output=65536.0)
    assert rel(dx1b / 65536.0, dx1a) < 1e-6 and rel(dx2b / 65536.0, dx2a) < 1e-6
    assert dlsb / 65536.0 == pytest.approx(dlsa, rel=1e-5)
    # d/dx of a function of x/||x|| is orthogonal to x
    cos = (dx1a * x1n).sum(-1) / (np.linalg.norm(dx1a, axis=-1) * np.linalg.norm(x1n, axis=-1))
    assert np.abs(cos).max() < 1e-3


def test_deterministic():
    x1n, x2n = io.make_pair(2000, 512, 0.3, 3)
    a = run(x1n, x2n, "bf16")
    b = run(x1n, x2n, "bf16")
    assert a[0] == b[0] and a[3] == b[3] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


def test_full_size_against_eager_formula():
    """BASELINE.json config: B = 32768, D = 512, bf16 mode, against the reference's formula
    (loss_head.py:271-283) executed by PyTorch eager in fp32 on the same GPU (TF32 off)."""
    B, D = 32768, 512
    torch.backends.cuda.matmul.allow_tf32 = False
    x1n, x2n = io.make_pair(B, D, 0.3, 1213)
    loss, dx1, dx2, dls = run(x1n, x2n, "bf16")
    x1 = torch.from_numpy(x1n).cuda().requires_grad_(True)
    x2 = torch.from_numpy(x2n).cuda().requires_grad_(True)
    ls = torch.tensor(math.log(1 / 0.07), device="cuda", requires_grad=True)
    a = x1 / x1.norm(dim=-1, keepdim=True)
    t = x2 / x2.norm(dim=-1, keepdim=True)
    s = ls.exp()
    labels = torch.arange(B, device="cuda")
    l12 = (s * a) @ t.t()
    ref = torch.nn.functional.cross_entropy(l12, labels)
    ref = ref + torch.nn.functional.cross_entropy(l12.t(), labels)      # (s*t)@a.T is l12.T up to rounding
    ref.backward()
    assert abs(loss - ref.item()) <= 1e-3 * abs(ref.item())
    assert rel(dx1, x1.grad.cpu().numpy()) <= 1e-2 and rel(dx2, x2.grad.cpu().numpy()) <= 1e-2
    assert abs(dls - ls.grad.item()) <= 1e-2 * max(abs(ls.grad.item()), 1e-3)


def test_full_size_invariants():
    """B = 32768 x 512 (BASELINE.json's size) through size-independent properties of the symmetric InfoNCE:
    swapping the modalities keeps the loss and swaps the gradients; permuting the pairs jointly permutes the gradients;
    every gradient row is orthogonal to its input row (normalisation Jacobian); rescaling a modality leaves the loss
    unchanged and divides its gradient."""
    B, D = 32768, 512
    x1n, x2n = io.make_pair(B, D, 0.3, 99)
    loss, dx1, dx2, dls = run(x1n, x2n, "bf16")
    loss_s, dx1_s, dx2_s, dls_s = run(x2n, x1n, "bf16")
    assert abs(loss - loss_s) <= 1e-5 * abs(loss) and abs(dls - dls_s) <= 1e-3 * abs(dls)
    assert rel(dx1_s, dx2) <= 1e-3 and rel(dx2_s, dx1) <= 1e-3
    perm = np.random.default_rng(5).permutation(B)
    loss_p, dx1_p, dx2_p, dls_p = run(np.ascontiguousarray(x1n[perm]), np.ascontiguousarray(x2n[perm]), "bf16")
    assert abs(loss - loss_p) <= 1e-5 * abs(loss) and abs(dls - dls_p) <= 1e-3 * abs(dls)
    assert rel(dx1_p, dx1[perm]) <= 1e-3 and rel(dx2_p, dx2[perm]) <= 1e-3
    for dx, x in ((dx1, x1n), (dx2, x2n)):
        dots = np.abs(np.einsum("ij,ij->i", dx.astype(np.float64), x.astype(np.float64)))
        bound = np.linalg.norm(dx, axis=1).astype(np.float64) * np.linalg.norm(x, axis=1)
        assert np.all(dots <= 1e-4 * bound + 1e-12)
    loss_c, dx1_c, _, _ = run(np.ascontiguousarray(4.0 * x1n), x2n, "bf16")      # power of two: the bf16 operands are identical
    assert abs(loss - loss_c) <= 1e-6 * abs(loss) and rel(4.0 * dx1_c, dx1) <= 1e-5


def test_row_shard_offsets_single_gpu():
    """The multi-GPU decomposition on one device: four row shards against the gathered matrices reproduce the
    single-shot statistics and gradients (C-ABI row_offset / rows_local / rows_global contract)."""
    from vipant_b200 import _cabi, functional as F_
    B, D, R = 1024, 512, 4
    x1n, x2n = io.make_pair(B, D, 0.3, 8)
    full = io.infonce_closed_form(x1n, x2n)
    x1, x2 = torch.from_numpy(x1n).cuda(), torch.from_numpy(x2n).cuda()
    ls = torch.tensor(math.log(1 / 0.07), device="cuda")
    K = F_._KERNELS
    prec = _cabi.PREC_BF16_TC
    a_all, t_all, inv, dcos = K.normalize_pair(x1, x2, False, prec)
    b = B // R
    stats, wss, scales, colsums = [], [], [], []
    for r in range(R):
        sl = slice(r * b, (r + 1) * b)
        cs, ws = K.forward_sweep(a_all[sl], t_all[sl], a_all, t_all, r * b, ls, None, prec)
        wss.append(ws); colsums.append(cs)
    col_sum = torch.stack(colsums).sum(0)                       # what the all-reduce produces
    for r in range(R):
        sl = slice(r * b, (r + 1) * b)
        st, sc = K.forward_finish(b, B, D, r * b, ls, None, dcos[sl], prec, wss[r], col_sum)
        stats.append(st); scales.append(sc)
    stats_all = torch.cat(stats, dim=1).contiguous()
    loss = K.loss(stats_all).item()
    assert abs(loss - full.loss) <= 1e-3 * abs(full.loss)
    np.testing.assert_allclose(stats_all[0].cpu().numpy(), full.row_lse, rtol=2e-3, atol=2e-3)
    np.testing.assert_allclose(stats_all[1].cpu().numpy(), full.col_lse, rtol=2e-3, atol=2e-3)
    dls_sum = 0.0
    for r in range(R):
        sl = slice(r * b, (r + 1) * b)
        dx1, dx2, dls = K.backward(x1[sl], x2[sl], a_all[sl], t_all[sl], a_all, t_all, inv[:, sl].contiguous(), stats_all,
                                   scales[r], wss[r], r * b, torch.tensor(1.0), False, prec)
        assert rel(dx1.cpu().numpy(), full.dx1[sl]) <= 1e-2 and rel(dx2.cpu().numpy(), full.dx2[sl]) <= 1e-2
        dls_sum += dls.item()
    assert abs(dls_sum - full.dlogit_scale) <= 1e-2 * abs(full.dlogit_scale)


def test_host_buffer_entry_point():
    """vpa_infonce_step_host: HOST pointers in, loss / gradients out (the call bench.py times as e2e)."""
    from vipant_b200 import _cabi
    lib = _cabi.lib()
    B, D = 640, 512
    x1n, x2n = io.make_pair(B, D, 0.3, 1213)
    ref = io.infonce_closed_form(x1n, x2n, grad_output=2.0)
    for prec, tol in ((_cabi.PREC_BF16_TC, TOL["bf16"]), (_cabi.PREC_FP32_SIMT, TOL["fp32"])):
        nbytes = lib.vpa_infonce_host_scratch_bytes(B, D, prec)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        loss, dls = ctypes.c_float(), ctypes.c_float()
        dx1, dx2 = np.empty_like(x1n), np.empty_like(x2n)
        rc = lib.vpa_infonce_step_host(x1n.ctypes.data, x2n.ctypes.data, B, D, math.log(1 / 0.07), 0.0, 2.0, prec,
                                       scratch.data_ptr(), nbytes, ctypes.addressof(loss), ctypes.addressof(dls),
                                       dx1.ctypes.data, dx2.ctypes.data, None)
        assert rc == 0, lib.vpa_last_error_string()
        assert abs(loss.value - ref.loss) <= tol["loss"] * abs(ref.loss)
        assert rel(dx1, ref.dx1) <= tol["grad"] and rel(dx2, ref.dx2) <= tol["grad"]
        assert abs(dls.value - ref.dlogit_scale) <= tol["grad"] * abs(ref.dlogit_scale)


@pytest.mark.parametrize("B,shards,ls,rho", [(2048, 4, math.log(1 / 0.07), 0.3), (1536, 3, math.log(1 / 0.07), 0.3),
                                             (1024, 2, math.log(100.0), 0.1)])
def test_host_buffer_entry_point_pipelined(B, shards, ls, rho, monkeypatch):
    """The pipelined host step: V row shards, copy-in / sweeps / copy-out overlapped on three streams (the default for
    batches >= 8192; forced here at a size the oracle handles).  ls = log(100) exercises the exact two-sweep regime."""
    from vipant_b200 import _cabi
    lib = _cabi.lib()
    D = 512
    monkeypatch.setenv("VPA_HOST_SHARDS", str(shards))
    x1n, x2n = io.make_pair(B, D, rho, 77)       # rho = 0.1 at s = 100: a loss that is not vanishingly small
    ref = io.infonce_closed_form(x1n, x2n, float(np.float32(ls)), None, False, 2.0)
    prec, tol = _cabi.PREC_BF16_TC, TOL["bf16"]
    nbytes = lib.vpa_infonce_host_scratch_bytes(B, D, prec)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    x1p, x2p = torch.from_numpy(x1n).pin_memory(), torch.from_numpy(x2n).pin_memory()
    dx1, dx2 = torch.empty_like(x1p).pin_memory(), torch.empty_like(x2p).pin_memory()
    for _ in range(2):            # twice: events / streams are reused across calls
        dx1.zero_(); dx2.zero_()
        loss, dls = ctypes.c_float(), ctypes.c_float()
        rc = lib.vpa_infonce_step_host(x1p.data_ptr(), x2p.data_ptr(), B, D, ls, 0.0, 2.0, prec, scratch.data_ptr(), nbytes,
                                       ctypes.addressof(loss), ctypes.addressof(dls), dx1.data_ptr(), dx2.data_ptr(),
                                       torch.cuda.current_stream().cuda_stream)
        assert rc == 0, lib.vpa_last_error_string()
        assert abs(loss.value - ref.loss) <= tol["loss"] * abs(ref.loss)
        assert rel(dx1.numpy(), ref.dx1) <= tol["grad"] and rel(dx2.numpy(), ref.dx2) <= tol["grad"]
        assert abs(dls.value - ref.dlogit_scale) <= tol["grad"] * abs(ref.dlogit_scale)


def test_loss_head_module_train_step_with_amp():
    """The monitors' step: autocast + GradScaler (cvap/monitor/cvap.py:189-193) through the drop-in module."""
    import vipant_b200 as vb
    B, D = 512, 512
    x1n, x2n = io.make_pair(B, D, 0.3, 1213)
    head = vb.build_loss_head(Cfg(name="CELossHead", scaling=True, scale_max=100.0)).cuda().train()
    x1 = torch.from_numpy(x1n).cuda().requires_grad_(True)
    x2 = torch.from_numpy(x2n).cuda().requires_grad_(True)
    scaler = torch.amp.GradScaler("cuda", init_scale=65536.0)
    with torch.autocast("cuda", dtype=torch.float16):
        loss = head(x1, x2, None, normalized=False, names=None)        # extra positional swallowed (cvalp.py:61)
    assert loss.dim() == 0 and loss.dtype == torch.float32
    scaler.scale(loss).backward()
    ref = io.infonce_closed_form(x1n, x2n, float(np.float32(np.log(1 / 0.07))), 100.0, False, 65536.0)
    assert abs(loss.item() - ref.loss) <= 1e-3 * abs(ref.loss)
    assert rel(x1.grad.cpu().numpy(), ref.dx1) <= 1e-2
    assert abs(head.logit_scale.grad.item() - ref.dlogit_scale) <= 1e-2 * abs(ref.dlogit_scale)
    assert head.logit_scale.grad.shape == ()


def test_composite_head_al_pair():
    """AT fine-tuning enters through VALCELossHead with only the `al` pair on (configs/model/loss/ce_val.yaml)."""
    import vipant_b200 as vb
    x1n, x2n = io.make_pair(64, 512, 0.3, 1213)
    head = vb.build_loss_head(Cfg(name="VALCELossHead", scaling=True, scale_max=None, va=False, lv=False, al=True)).cuda().train()
    aud = torch.from_numpy(x1n).cuda().requires_grad_(True)
    txt = torch.from_numpy(x2n).cuda().requires_grad_(True)
    loss = head(None, aud, txt, normalized=False, names=None)
    loss.backward()
    g = load_golden("infonce_c1_b64")
    assert abs(loss.item() - float(g["loss"])) <= 1e-3 * float(g["loss"])
    assert rel(aud.grad.cpu().numpy(), g["dx1"]) <= 1e-2 and rel(txt.grad.cpu().numpy(), g["dx2"]) <= 1e-2
    assert head.stats(nstep=1).startswith("al ")


def test_row_strided_views_get_correct_gradients():
    """Inputs that are row-strided views (`feats[:, :512]`, the CLS token `hidden[:, 0, :]`) are read in place; their
    gradient buffers must carry the same leading dimension (`empty_like` densifies a view and the finalize
    kernel addresses x and dx with one ld -> out-of-bounds writes)."""
    import vipant_b200 as vb
    B, D = 384, 512
    x1n, x2n = io.make_pair(B, D, 0.3, 77)
    ref = io.infonce_closed_form(x1n, x2n)
    for precision in ("bf16", "fp32"):
        wide = torch.zeros(B, 768, device="cuda")
        wide[:, :D] = torch.from_numpy(x1n).cuda()
        hidden = torch.zeros(B, 3, D, device="cuda")
        hidden[:, 0, :] = torch.from_numpy(x2n).cuda()
        guard1, guard2 = wide[:, D:].clone(), hidden[:, 1:, :].clone()
        wide.requires_grad_(True)
        hidden.requires_grad_(True)
        v1, v2 = wide[:, :D], hidden[:, 0, :]
        assert v1.stride(0) == 768 and v2.stride(0) == 3 * D
        ls = torch.tensor(math.log(1 / 0.07), device="cuda", requires_grad=True)
        loss = vb.infonce_loss(v1, v2, ls, precision=precision)
        loss.backward()
        tol = TOL[precision]
        assert abs(loss.item() - ref.loss) <= tol["loss"] * abs(ref.loss)
        assert rel(wide.grad[:, :D].cpu().numpy(), ref.dx1) <= tol["grad"]
        assert rel(hidden.grad[:, 0, :].cpu().numpy(), ref.dx2) <= tol["grad"]
        assert float(wide.grad[:, D:].abs().max()) == 0.0 and float(hidden.grad[:, 1:, :].abs().max()) == 0.0
        assert torch.equal(wide.detach()[:, D:], guard1) and torch.equal(hidden.detach()[:, 1:, :], guard2)


def test_constant_temperature_head_scaling_false():
    """`scaling=False` (loss_head.py:252): logit_scale is a plain CPU tensor log(1) = 0, never moved by .cuda(), no gradient."""
    import vipant_b200 as vb
    x1n, x2n = io.make_pair(200, 512, 0.3, 5)
    head = vb.build_loss_head(Cfg(name="CELossHead", scaling=False, scale_max=None)).cuda().train()
    assert not isinstance(head.logit_scale, torch.nn.Parameter) and head.logit_scale.device.type == "cpu"
    x1 = torch.from_numpy(x1n).cuda().requires_grad_(True)
    x2 = torch.from_numpy(x2n).cuda().requires_grad_(True)
    for _ in range(2):              # second call: the cached device copy of the constant
        x1.grad = x2.grad = None
        loss = head(x1, x2)
        loss.backward()
    ref = io.infonce_closed_form(x1n, x2n, 0.0)
    assert abs(loss.item() - ref.loss) <= 1e-3 * abs(ref.loss)
    assert rel(x1.grad.cpu().numpy(), ref.dx1) <= 1e-2 and rel(x2.grad.cpu().numpy(), ref.dx2) <= 1e-2


def test_vace_head_forward_weighted_pairs():
    """VACELossHead.forward (loss_head.py:497-598): five weighted InfoNCE pairs over five feature matrices, each pair with
    its own temperature; gradients of a shared modality add up over its pairs."""
    import vipant_b200 as vb
    B, D = 256, 512
    rng = np.random.default_rng(3)
    feats = [rng.standard_normal((B, D)).astype(np.float32) for _ in range(5)]      # images, images_v1, audios_v1, images_v2, audios_v2
    w = dict(vp_w=1.0, ap_w=0.5, va_w=2.0, vv_w=0.25, aa_w=0.75)
    head = vb.build_loss_head(Cfg(name="VACELossHead", scaling=True, scale_max=None, vp=True, ap=True, va=True, vv=True, aa=True,
                                  **w)).cuda().train()
    xs = [torch.from_numpy(f).cuda().requires_grad_(True) for f in feats]
    loss = head(*xs, normalized=False)
    loss.backward()
    pairs = [(1, 0, w["vp_w"]), (2, 0, w["ap_w"]), (1, 2, w["va_w"]), (1, 3, w["vv_w"]), (2, 4, w["aa_w"])]
    want = 0.0
    grads = [np.zeros((B, D)) for _ in range(5)]
    for i, j, wt in pairs:
        r = io.infonce_closed_form(feats[i], feats[j], grad_output=wt)
        want += wt * r.loss
        grads[i] += r.dx1
        grads[j] += r.dx2
    assert abs(loss.item() - want) <= 1e-3 * abs(want)
    for x, g in zip(xs, grads):
        assert rel(x.grad.cpu().numpy(), g) <= 1e-2
    assert head.stats(nstep=1).split()[0::2] == ["vp", "ap", "va", "vv", "aa"]


def test_one_process_two_devices():
    """The reference's `dp` mode is one process driving several GPUs: per-device kernel attributes (232 KB dynamic shared
    memory) and the SM-count cache must follow the current device."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import vipant_b200 as vb
    x1n, x2n = io.make_pair(512, 512, 0.3, 9)
    ref = io.infonce_closed_form(x1n, x2n)
    for dev in (0, 1, 0):
        x1 = torch.from_numpy(x1n).to(f"cuda:{dev}").requires_grad_(True)
        x2 = torch.from_numpy(x2n).to(f"cuda:{dev}").requires_grad_(True)
        ls = torch.tensor(math.log(1 / 0.07), device=f"cuda:{dev}", requires_grad=True)
        loss = vb.infonce_loss(x1, x2, ls, precision="bf16")
        loss.backward()
        assert abs(loss.item() - ref.loss) <= 1e-3 * abs(ref.loss) and rel(x1.grad.cpu().numpy(), ref.dx1) <= 1e-2
