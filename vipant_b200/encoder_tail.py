"""Encoder tail fused up to the InfoNCE operand.

The reference's towers end with  ``x = self.ln(x[:, 0, :]); x = x @ self.proj``  (``ViTPostEncoder``,
``/root/reference/cvap/module/val.py:288-290``; ``GPTPostEncoder`` :143-146 picks the EOT token instead of the CLS token) and
the head then normalises  ``x = x / x.norm(dim=-1, keepdim=True)``  when the loss head asks for it
(``cvap/module/encoder/clip_head.py:117-118``, ``audio_head.py:209-210``).  ``vpa_encoder_tail`` does all three in two kernels
(LayerNorm + cast; tensor-core projection whose epilogue normalises and emits the bf16 operand rows), and this module wraps it
behind the reference's post-encoder interface: same parameter names (``ln.weight``, ``ln.bias``, ``proj``), same forward
signature, same ``normalized=`` keyword as the heads use.

  * ``forward(x)`` returns the projected features like the reference does (fp32; ``normalized=True``: unit rows).
  * ``operands(x)`` returns ``(a_bf16, inv_norm, y)``: the normalised bf16 rows in the layout the sweep kernels read.
  * Training: the backward of the tail is two plain GEMMs (``dproj = ln^T dy``, ``dln = dy proj^T`` -- library GEMMs, like the
    reference's own autograd) plus the LayerNorm backward; the forward kept the LayerNorm output and statistics for it.

CUDA only, like everything in this package: CPU tensors raise.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _cabi
from . import functional as F_

__all__ = ["encoder_tail", "FusedPostEncoder"]


def _run_tail(x, gamma, beta, proj_t_bf16, eps, want_y, want_stats):
    lib = _cabi.lib()
    rows, width = x.shape
    N = proj_t_bf16.shape[0]
    dev = x.device
    ln = torch.empty((rows, width), dtype=torch.bfloat16, device=dev)
    a = torch.empty((rows, N), dtype=torch.bfloat16, device=dev)
    inv = torch.empty((rows,), dtype=torch.float32, device=dev)
    y = torch.empty((rows, N), dtype=torch.float32, device=dev) if want_y else None
    stats = torch.empty((2, rows), dtype=torch.float32, device=dev) if want_stats else None
    with torch.cuda.device(dev):
        _cabi.check(lib.vpa_encoder_tail(
            x.data_ptr(), F_._DTYPES[x.dtype], rows, width, x.stride(0), gamma.data_ptr(), beta.data_ptr(), float(eps),
            proj_t_bf16.data_ptr(), N, ln.data_ptr(), None if stats is None else stats[0].data_ptr(),
            None if stats is None else stats[1].data_ptr(), a.data_ptr(), None if y is None else y.data_ptr(), inv.data_ptr(),
            torch.cuda.current_stream().cuda_stream), "vpa_encoder_tail")
    return a, inv, y, ln, stats


def transposed_bf16(proj):
    """(embed_dim, width) bf16, K-major: what the TMA map of the projection (the B operand) reads."""
    return proj.detach().t().contiguous().to(torch.bfloat16)


def _prep(x, gamma, beta, proj, proj_t_bf16=None):
    F_._require_cuda(x, gamma, beta, proj)
    x = F_._rows2d(x)                                  # row-strided CLS views are read in place
    if proj.dim() != 2 or proj.shape[0] != x.shape[1]:
        raise ValueError(f"proj must be (width={x.shape[1]}, embed_dim), got {tuple(proj.shape)}")
    g = gamma.detach().float().contiguous()
    b = beta.detach().float().contiguous()
    pt = transposed_bf16(proj) if proj_t_bf16 is None else proj_t_bf16
    if pt.dtype != torch.bfloat16 or tuple(pt.shape) != (proj.shape[1], proj.shape[0]) or not pt.is_contiguous():
        raise ValueError("proj_t_bf16 must be the contiguous bf16 transpose of proj")
    return x, g, b, pt


class _TailFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, proj, eps, proj_t_bf16):
        xc, g, b, pt = _prep(x, gamma, beta, proj, proj_t_bf16)
        a, inv, y, ln, stats = _run_tail(xc, g, b, pt, eps, True, True)
        ctx.save_for_backward(xc, gamma, proj, ln, stats)
        ctx.mark_non_differentiable(a, inv)
        return y, a, inv

    @staticmethod
    def backward(ctx, dy, _da, _dinv):
        xc, gamma, proj, ln, stats = ctx.saved_tensors
        return tail_backward(xc, gamma, proj, ln, stats[0], stats[1], dy) + (None, None)


def tail_backward(x, gamma, proj, ln, mean, rstd, dy):
    """Gradients of the tail w.r.t. (x, ln.weight, ln.bias, proj) from d/dy: two plain library GEMMs (as the reference's autograd
    runs them) and the LayerNorm backward from the statistics the forward kept.  ln: the forward's bf16 LayerNorm output."""
    dy = dy.float()
    dproj = (ln.float().t() @ dy).to(proj.dtype)
    dln = dy @ proj.detach().float().t()
    mean, rstd = mean[:, None], rstd[:, None]
    xhat = (x.detach().float() - mean) * rstd
    dgamma = (dln * xhat).sum(0).to(gamma.dtype)
    dbeta = dln.sum(0).to(gamma.dtype)
    dxh = dln * gamma.detach().float()
    dx = rstd * (dxh - dxh.mean(-1, keepdim=True) - xhat * (dxh * xhat).mean(-1, keepdim=True))
    return dx.to(x.dtype), dgamma, dbeta, dproj


def encoder_tail(x, gamma, beta, proj, eps: float = 1e-5, need_grad: bool = None, proj_t_bf16=None, features: bool = True):
    """LayerNorm(x) @ proj, L2-normalised and cast: returns ``(y, a_bf16, inv_norm)`` -- y (rows, N) fp32 un-normalised features
    (differentiable), a_bf16 the unit rows as bf16 operands, inv_norm = 1 / ||y||.  x: (rows, width) CUDA, any row stride.
    proj_t_bf16: ``transposed_bf16(proj)`` computed by the caller (the module keeps it per parameter version); by default it
    is made here, two small launches per call.  features=False (inference only): y is not written and comes back as None."""
    if need_grad is None:
        need_grad = torch.is_grad_enabled() and any(t.requires_grad for t in (x, gamma, beta, proj))
    if need_grad:
        return _TailFunction.apply(x, gamma, beta, proj, eps, proj_t_bf16)
    xc, g, b, pt = _prep(x, gamma, beta, proj, proj_t_bf16)
    a, inv, y, _, _ = _run_tail(xc, g, b, pt, eps, bool(features), False)
    return y, a, inv


class FusedPostEncoder(nn.Module):
    """Drop-in for the reference's ``ViTPostEncoder`` / ``GPTPostEncoder`` tails (val.py:262-290, :125-146): parameters ``ln``
    (weight, bias) and ``proj`` under the same names, ``forward(x, mask=None, normalized=False)``.

    x: (batch, tokens, width) hidden states -- the CLS token ``x[:, 0, :]`` (or, with ``mask`` = EOT positions, the rows
    ``x[arange, mask]``) goes through the fused kernels -- or an already selected (batch, width) matrix.
    """

    def __init__(self, cfg=None, width: int = 768, embed_dim: int = 512, **kwargs):
        super().__init__()
        self.ln = nn.LayerNorm(width)
        self.proj = nn.Parameter(width ** -0.5 * torch.randn(width, embed_dim))
        self._proj_t = None                            # (key, bf16 transpose): redone when the parameter changes

    def _proj_operand(self):
        p = self.proj
        key = (p.data_ptr(), p._version, p.device)     # in-place optimiser steps / load_state_dict bump _version
        if self._proj_t is None or self._proj_t[0] != key:
            self._proj_t = (key, transposed_bf16(p))
        return self._proj_t[1]

    def _tail(self, x, mask, features=True):
        return encoder_tail(self._select(x, mask), self.ln.weight, self.ln.bias, self.proj, self.ln.eps,
                            proj_t_bf16=self._proj_operand(), features=features)

    def _select(self, x, mask):
        if x.dim() == 2:
            return x
        if mask is None:
            return x[:, 0, :]
        return x[torch.arange(x.shape[0], device=x.device), mask]

    def operands(self, x, mask=None, features: bool = True):
        """(a_bf16, inv_norm, y): the unit rows as the sweep kernels read them, 1/||y||, and the raw projected features
        (None with features=False under no_grad: the fp32 features are then never written)."""
        y, a, inv = self._tail(x, mask, features)
        return a, inv, y

    def forward(self, x, positional_embedding=None, class_embedding=None, mask=None, normalized: bool = False, **kwargs):
        y, a, inv = self._tail(x, mask)
        if normalized:
            if y.requires_grad:                        # training: the normalisation's Jacobian belongs to autograd
                return y / y.norm(dim=-1, keepdim=True)
            return y * inv[:, None]
        return y
