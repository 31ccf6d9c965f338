"""Encoder tail (LayerNorm -> projection -> L2 normalise -> bf16 operand rows) against a plain PyTorch fp32 restatement of the
reference's three lines (cvap/module/val.py:288-290 `x = self.ln(x[:, 0, :]); x = x @ self.proj`, clip_head.py:117-118
`x = x / x.norm(dim=-1, keepdim=True)`).

Floating point.  The kernel rounds the LayerNorm output and the projection to bf16 and accumulates in fp32 on the tensor cores:
  * against the SAME rounding done in torch (bf16 operands, fp32 matmul) the features agree to 2e-3 of a row's norm;
  * against the pure fp32 restatement the tolerance is the bf16 operand rounding: 1.5e-2 of a row's norm;
  * LayerNorm statistics, 1/||y|| and the emitted bf16 operand rows are tight (1e-5 relative / one bf16 ulp).
"""
import pytest
import torch
import torch.nn.functional as TF

pytestmark = pytest.mark.gpu

from vipant_b200.encoder_tail import FusedPostEncoder, encoder_tail  # noqa: E402


def _ref(x, gamma, beta, proj, eps=1e-5, round_bf16=False):
    ln = TF.layer_norm(x.float(), (x.shape[-1],), gamma.float(), beta.float(), eps)
    p = proj.float()
    if round_bf16:
        ln = ln.to(torch.bfloat16).float()
        p = p.to(torch.bfloat16).float()
    y = ln @ p
    return ln, y


def _mk(rows, width, N, seed, dtype=torch.float32):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = (torch.randn(rows, width, device="cuda", generator=g) * 1.7 + 0.3).to(dtype)
    gamma = 1.0 + 0.2 * torch.randn(width, device="cuda", generator=g)
    beta = 0.1 * torch.randn(width, device="cuda", generator=g)
    proj = width ** -0.5 * torch.randn(width, N, device="cuda", generator=g)
    return x, gamma, beta, proj


def _check_forward(x, gamma, beta, proj):
    torch.backends.cuda.matmul.allow_tf32 = False
    y, a, inv = encoder_tail(x, gamma, beta, proj, need_grad=False)
    torch.cuda.synchronize()
    _, y_r = _ref(x, gamma, beta, proj, round_bf16=True)
    _, y_f = _ref(x, gamma, beta, proj)
    norm = y_f.norm(dim=-1, keepdim=True)
    assert y.shape == y_f.shape and y.dtype == torch.float32
    assert ((y - y_r).abs().max(dim=-1, keepdim=True).values / norm).max().item() < 2e-3
    assert ((y - y_f).norm(dim=-1, keepdim=True) / norm).max().item() < 1.5e-2
    torch.testing.assert_close(inv, 1.0 / y.norm(dim=-1), rtol=1e-5, atol=0)
    unit = (y * inv[:, None])
    assert a.dtype == torch.bfloat16
    # one bf16 ulp around the fp32 unit rows (values < 1: ulp <= 2^-8 relative)
    assert ((a.float() - unit).abs() <= unit.abs() * 2 ** -8 + 1e-30).all()
    return y, a, inv


@pytest.mark.parametrize("rows,width,N", [(1, 768, 512), (100, 768, 512), (128, 512, 512), (333, 1024, 512), (1024, 768, 256),
                                           (4096, 768, 512), (257, 64, 256)])
def test_forward_matches_fp32_reference(rows, width, N):
    _check_forward(*_mk(rows, width, N, seed=rows + width))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_half_inputs(dtype):
    _check_forward(*_mk(300, 768, 512, seed=7, dtype=dtype))


def test_cls_token_view_is_read_in_place():
    g = torch.Generator(device="cuda").manual_seed(3)
    hidden = torch.randn(200, 5, 768, device="cuda", generator=g)
    _, gamma, beta, proj = _mk(1, 768, 512, seed=4)
    cls = hidden[:, 0, :]
    assert not cls.is_contiguous()
    y, a, inv = _check_forward(cls, gamma, beta, proj)
    y2, a2, inv2 = encoder_tail(cls.contiguous(), gamma, beta, proj, need_grad=False)
    assert torch.equal(y, y2) and torch.equal(a, a2) and torch.equal(inv, inv2)


def test_layernorm_statistics_and_determinism():
    from vipant_b200.encoder_tail import _prep, _run_tail
    x, gamma, beta, proj = _mk(500, 768, 512, seed=11)
    xc, g, b, pt = _prep(x, gamma, beta, proj)
    a, inv, y, ln, stats = _run_tail(xc, g, b, pt, 1e-5, True, True)
    ln_r, _ = _ref(x, gamma, beta, proj)
    torch.testing.assert_close(stats[0], x.mean(-1), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(stats[1], torch.rsqrt(x.var(-1, unbiased=False) + 1e-5), rtol=1e-5, atol=0)
    assert ((ln.float() - ln_r).abs() <= ln_r.abs() * 2 ** -8 + 1e-6).all()
    a2, inv2, y2, _, _ = _run_tail(xc, g, b, pt, 1e-5, True, False)
    assert torch.equal(a, a2) and torch.equal(inv, inv2) and torch.equal(y, y2)


def test_gradients_match_autograd_of_the_reference():
    x, gamma, beta, proj = _mk(384, 768, 512, seed=21)
    w = torch.randn(384, 512, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
    leaves = [t.clone().requires_grad_(True) for t in (x, gamma, beta, proj)]
    y, a, inv = encoder_tail(*leaves)
    assert y.requires_grad and not a.requires_grad and not inv.requires_grad
    # a loss through the normalised rows, like the head's:  sum(w * y / ||y||)
    ((y / y.norm(dim=-1, keepdim=True)) * w).sum().backward()
    ref = [t.clone().requires_grad_(True) for t in (x, gamma, beta, proj)]
    _, y_r = _ref(*ref)
    ((y_r / y_r.norm(dim=-1, keepdim=True)) * w).sum().backward()
    for got, want, name in zip(leaves, ref, ("x", "ln.weight", "ln.bias", "proj")):
        rel = (got.grad - want.grad).norm() / want.grad.norm()
        assert rel.item() < 1.5e-2, (name, rel.item())


def test_module_mirrors_the_reference_post_encoder():
    torch.manual_seed(0)
    m = FusedPostEncoder(None, width=768, embed_dim=512).cuda()
    assert sorted(m.state_dict()) == ["ln.bias", "ln.weight", "proj"]      # the reference's checkpoint keys (val.py:270-273)
    hidden = torch.randn(64, 7, 768, device="cuda")
    out = m(hidden, normalized=True)
    _, y_f = _ref(hidden[:, 0, :], m.ln.weight, m.ln.bias, m.proj)
    want = y_f / y_f.norm(dim=-1, keepdim=True)
    assert (out - want).norm(dim=-1).max().item() < 1.5e-2
    # EOT selection (GPTPostEncoder, val.py:143-146)
    eot = torch.randint(0, 7, (64,), device="cuda")
    out = m(hidden, mask=eot)
    _, y_f = _ref(hidden[torch.arange(64, device="cuda"), eot], m.ln.weight, m.ln.bias, m.proj)
    assert ((out - y_f).norm(dim=-1) / y_f.norm(dim=-1)).max().item() < 1.5e-2
    out.sum().backward()
    assert m.proj.grad is not None and m.ln.weight.grad is not None


def test_operands_feed_the_loss_head():
    """The emitted bf16 rows are unit rows in the layout the loss reads: the InfoNCE loss on them (normalized=True) equals the
    closed form on the fp32 restatement's unit rows within the bf16 tolerance of the loss tests (2e-2 absolute at scale 14)."""
    import vipant_b200
    B = 512
    xa, ga, ba, pa = _mk(B, 768, 512, seed=31)
    xt, gt, bt, pt = _mk(B, 512, 512, seed=32)
    _, a, _ = encoder_tail(xa, ga, ba, pa, need_grad=False)
    _, t, _ = encoder_tail(xt, gt, bt, pt, need_grad=False)
    ls = torch.tensor(2.0, device="cuda")
    loss = vipant_b200.infonce_loss(a, t, ls, normalized=True)
    ya = _ref(xa, ga, ba, pa)[1].double()
    yt = _ref(xt, gt, bt, pt)[1].double()
    ya = ya / ya.norm(dim=-1, keepdim=True)
    yt = yt / yt.norm(dim=-1, keepdim=True)
    logits = ls.double().exp() * ya @ yt.t()
    lab = torch.arange(B, device="cuda")
    want = TF.cross_entropy(logits, lab) + TF.cross_entropy(logits.t(), lab)
    assert abs(loss.item() - want.item()) < 2e-2


def test_unsupported_shapes_fail_loudly():
    from vipant_b200._cabi import VipantB200Error
    x, gamma, beta, proj = _mk(16, 768, 384, seed=1)
    with pytest.raises(VipantB200Error):
        encoder_tail(x, gamma, beta, proj, need_grad=False)
    x, gamma, beta, proj = _mk(16, 1088, 512, seed=1)
    with pytest.raises(VipantB200Error):
        encoder_tail(x, gamma, beta, proj, need_grad=False)


# ---------------------------------------------------------------- against the reference's own post-encoders (golden) and the oracle
GOLDEN_CASES = {"vit": dict(seed=4101, rows=72, tokens=3, width=768, embed=512),
                "gpt": dict(seed=4102, rows=40, tokens=6, width=512, embed=512),
                "vit256": dict(seed=4103, rows=130, tokens=2, width=1024, embed=256)}


@pytest.mark.parametrize("name", sorted(GOLDEN_CASES))
def test_module_matches_reference_post_encoder_golden(name):
    """FusedPostEncoder against tests/golden/encoder_tail.npz (outputs of the reference's ViTPostEncoder / GPTPostEncoder and the
    heads' normalisation, fp32 CPU; oracle/make_golden_encoder_tail.py).  Tolerance = bf16 operand rounding: 1.5e-2 of a row's
    norm for the features, 1.5e-2 absolute L2 distance for the unit rows; 1/||y|| within 1e-2 relative."""
    import numpy as np
    from conftest import load_golden
    from oracle import encoder_tail_oracle as eo
    c = GOLDEN_CASES[name]
    fx = load_golden("encoder_tail")
    inp = eo.golden_inputs(**c)
    m = FusedPostEncoder(None, width=c["width"], embed_dim=c["embed"]).cuda()
    with torch.no_grad():
        m.ln.weight.copy_(torch.from_numpy(inp["gamma"]))
        m.ln.bias.copy_(torch.from_numpy(inp["beta"]))
        m.proj.copy_(torch.from_numpy(inp["proj"]))
    hidden = torch.from_numpy(inp["hidden"]).cuda()
    mask = torch.from_numpy(inp["eot"]).cuda() if name == "gpt" else None
    with torch.no_grad():
        y = m(hidden, mask=mask)
        unit = m(hidden, mask=mask, normalized=True)
        a, inv, _ = m.operands(hidden, mask=mask)
    want_y = torch.from_numpy(fx[f"{name}_y"]).cuda()
    want_norm = torch.from_numpy(fx[f"{name}_norm"]).cuda()
    assert ((y - want_y).norm(dim=-1) / want_norm).max().item() < 1.5e-2
    assert ((unit - want_y / want_norm[:, None]).norm(dim=-1)).max().item() < 1.5e-2
    assert ((a.float() - want_y / want_norm[:, None]).norm(dim=-1)).max().item() < 2e-2
    assert (inv * want_norm - 1).abs().max().item() < 1e-2
    # and against the float64 oracle (same bound: the oracle agrees with the golden to 2e-5)
    y64, _ = eo.encoder_tail(inp["hidden"], inp["gamma"], inp["beta"], inp["proj"], mask=inp["eot"] if name == "gpt" else None)
    assert (np.linalg.norm(y.cpu().numpy() - y64, axis=-1) / np.linalg.norm(y64, axis=-1)).max() < 1.5e-2


def test_gradients_match_reference_autograd_golden():
    """Gradients of sum(w * y/||y||) through the fused tail against the reference autograd's (golden, ViT case)."""
    import numpy as np
    from conftest import load_golden
    from oracle import encoder_tail_oracle as eo
    from oracle.make_golden_encoder_tail import GRAD_ROWS
    c = GOLDEN_CASES["vit"]
    fx = load_golden("encoder_tail")
    inp = eo.golden_inputs(**c)
    m = FusedPostEncoder(None, width=c["width"], embed_dim=c["embed"]).cuda()
    with torch.no_grad():
        m.ln.weight.copy_(torch.from_numpy(inp["gamma"]))
        m.ln.bias.copy_(torch.from_numpy(inp["beta"]))
        m.proj.copy_(torch.from_numpy(inp["proj"]))
    hidden = torch.from_numpy(inp["hidden"]).cuda().requires_grad_(True)
    unit = m(hidden, normalized=True)
    (unit * torch.from_numpy(inp["w"]).cuda()).sum().backward()
    assert hidden.grad[:, 1:, :].abs().max().item() == 0.0
    dx = hidden.grad[:, 0, :].cpu().numpy()
    rows = np.asarray(GRAD_ROWS)

    def rel(got, want):
        return float(np.linalg.norm(got - want) / np.linalg.norm(want))
    assert rel(dx[rows], fx["vit_dx_rows"]) < 1.5e-2
    assert abs(np.linalg.norm(dx) / float(fx["vit_dx_norm"]) - 1) < 1e-2
    assert rel(m.ln.weight.grad.cpu().numpy(), fx["vit_dgamma"]) < 1.5e-2
    assert rel(m.ln.bias.grad.cpu().numpy(), fx["vit_dbeta"]) < 1.5e-2
    assert rel(m.proj.grad.cpu().numpy()[rows], fx["vit_dproj_rows"]) < 1.5e-2
    assert abs(np.linalg.norm(m.proj.grad.cpu().numpy()) / float(fx["vit_dproj_norm"]) - 1) < 1e-2


def test_tail_and_loss_chain_gradients():
    """Hidden states -> fused tails -> InfoNCE loss -> backward, against the same chain in PyTorch fp32 (the reference's
    arithmetic: LayerNorm, matmul, normalise, scaled logits, two cross entropies).  Tolerances: loss 1e-3 relative (the loss
    tests' bf16 bar), gradients 5e-2 relative -- a composition check (bf16 operands in the projection AND in the loss sweeps; each stage
    has its own tighter test above / in test_gpu_infonce.py)."""
    import vipant_b200
    B = 256
    xa, ga, ba, pa = _mk(B, 768, 512, seed=41)
    xt, gt, bt, pt = _mk(B, 512, 512, seed=42)
    xt = 0.6 * xt + 0.4 * xa[:, :512]                     # correlated pairs: a non-trivial diagonal
    ls0 = 2.3

    def run(fused):
        leaves = [t.clone().requires_grad_(True) for t in (xa, ga, ba, pa, xt, gt, bt, pt)]
        ls = torch.tensor(ls0, device="cuda", requires_grad=True)
        if fused:
            ya = encoder_tail(*leaves[:4])[0]
            yt = encoder_tail(*leaves[4:])[0]
            loss = vipant_b200.infonce_loss(ya, yt, ls)
        else:
            ya = _ref(*leaves[:4])[1]
            yt = _ref(*leaves[4:])[1]
            ya = ya / ya.norm(dim=-1, keepdim=True)
            yt = yt / yt.norm(dim=-1, keepdim=True)
            logits = ls.exp() * ya @ yt.t()
            lab = torch.arange(B, device="cuda")
            loss = TF.cross_entropy(logits, lab) + TF.cross_entropy(logits.t(), lab)
        loss.backward()
        return loss.item(), [t.grad for t in leaves] + [ls.grad]

    torch.backends.cuda.matmul.allow_tf32 = False
    loss_f, grads_f = run(True)
    loss_r, grads_r = run(False)
    assert abs(loss_f - loss_r) < 1e-3 * abs(loss_r) + 1e-3
    names = ("x_a", "ln_a.weight", "ln_a.bias", "proj_a", "x_t", "ln_t.weight", "ln_t.bias", "proj_t", "logit_scale")
    for got, want, name in zip(grads_f, grads_r, names):
        rel = ((got - want).norm() / want.norm()).item()
        assert rel < 5e-2, (name, rel)


def test_module_refreshes_its_projection_operand_when_the_parameter_changes():
    """The module keeps the transposed bf16 projection per parameter version: an optimiser step (in place), load_state_dict
    and a freshly assigned .data must all be picked up."""
    torch.manual_seed(1)
    m = FusedPostEncoder(None, width=512, embed_dim=256).cuda()
    x = torch.randn(40, 512, device="cuda")
    with torch.no_grad():
        y0 = m(x).clone()
        assert m._proj_t is not None
        a_only, inv_only, none = m.operands(x, features=False)   # inference: the fp32 features are not written
        a_full, inv_full, y_full = m.operands(x)
        assert none is None and torch.equal(a_only, a_full) and torch.equal(inv_only, inv_full) and torch.equal(y_full, y0)
        cached = m._proj_t[1]
        assert m._proj_operand() is cached                       # unchanged parameter: reused
        m.proj.mul_(2.0)                                         # in-place update
        y1 = m(x)
        torch.testing.assert_close(y1, 2.0 * y0, rtol=1e-6, atol=1e-6)      # exact: a power of two commutes with the bf16 rounding
        sd = {k: v.clone() for k, v in m.state_dict().items()}
        sd["proj"] = sd["proj"] * 0.5
        m.load_state_dict(sd)
        torch.testing.assert_close(m(x), y0, rtol=1e-6, atol=1e-6)
        m.proj.data = (4.0 * m.proj.data).clone()
        torch.testing.assert_close(m(x), 4.0 * y0, rtol=1e-6, atol=1e-6)
