"""CPU restatement of the reference's contrastive InfoNCE head (test infrastructure).

Follows ``/root/reference/cvap/module/decoder/loss_head.py``:
  * normalisation            -> loss_head.py:271-273   (``x / x.norm(dim=-1, keepdim=True)``, no eps)
  * temperature              -> loss_head.py:276       (``logit_scale.exp().clamp(max=scale_max)``)
  * logits, both directions  -> loss_head.py:277-278   (``(s * x1) @ x2.t()`` and its transpose)
  * 2 x cross-entropy, SUM   -> loss_head.py:280-283   (``nn.CrossEntropyLoss()`` mean reduction each)
  * ``scale_max or inf``     -> loss_head.py:254

Two restatements:
  ``infonce_closed_form``  numpy, any float dtype (fp64 by default): loss, the
      intermediate statistics (row/column logsumexp, diagonal) and the analytic
      gradients.  This is the checker for the CUDA path.
  ``infonce_port_torch``   the same arithmetic written with torch CPU ops and
      autograd, i.e. what the reference executes on a CPU; it is the timed
      ``cpu_baseline`` ("port") in bench.py and a second witness in the tests.
  ``infonce_row_sharded``  the row-sharded decomposition the multi-GPU path
      uses, restated on one process to check the sharding algebra.

Pinned against the reference's own code through tests/golden/*.npz
(oracle/make_golden.py); see tests/test_oracle.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np


@dataclass
class InfoNCEResult:
    loss: float
    row_lse: np.ndarray   # (B,) logsumexp_j S_ij
    col_lse: np.ndarray   # (B,) logsumexp_i S_ij
    diag: np.ndarray      # (B,) S_ii
    scale: float          # s = min(exp(logit_scale), scale_max)
    dx1: np.ndarray       # dL/dx1 (B, D)
    dx2: np.ndarray       # dL/dx2 (B, D)
    dlogit_scale: float   # dL/d logit_scale (0 when the clamp saturates)


def _lse(S, axis):
    m = S.max(axis=axis, keepdims=True)
    return (m + np.log(np.exp(S - m).sum(axis=axis, keepdims=True))).squeeze(axis)


def effective_scale(logit_scale: float, scale_max) -> tuple[float, bool]:
    """loss_head.py:254,276 -- returns (s, gradient_flows)."""
    cap = scale_max or float("inf")       # `cfg.scale_max or float("inf")`: 0/None -> inf
    e = math.exp(logit_scale)
    # torch.clamp(max=) passes the gradient where input <= max
    return (min(e, cap), e <= cap)


def infonce_closed_form(x1, x2, logit_scale=math.log(1 / 0.07), scale_max=None,
                        normalized=False, grad_output=1.0, dtype=np.float64) -> InfoNCEResult:
    x1 = np.asarray(x1, dtype=dtype)
    x2 = np.asarray(x2, dtype=dtype)
    B = x1.shape[0]
    if not normalized:                    # loss_head.py:271-273
        n1 = np.sqrt((x1 * x1).sum(-1, keepdims=True))
        n2 = np.sqrt((x2 * x2).sum(-1, keepdims=True))
        a, t = x1 / n1, x2 / n2
    else:
        a, t = x1, x2
    s, flows = effective_scale(float(logit_scale), scale_max)
    S = (dtype(s) * a) @ t.T              # loss_head.py:277 (278 is its transpose)
    row_lse = _lse(S, 1)
    col_lse = _lse(S, 0)
    diag = np.diagonal(S).copy()
    loss = (row_lse - diag).mean() + (col_lse - diag).mean()   # loss_head.py:280-283
    # dL/dS = (softmax_rows + softmax_cols - 2 I) / B
    G = np.exp(S - row_lse[:, None]) + np.exp(S - col_lse[None, :])
    G[np.arange(B), np.arange(B)] -= 2.0
    G *= dtype(grad_output) / B
    da = dtype(s) * (G @ t)
    dt = dtype(s) * (G.T @ a)
    dls = float((G * S).sum()) if flows else 0.0      # d/dl of s*cos = S (s = exp(l))
    if not normalized:
        dx1 = (da - a * (a * da).sum(-1, keepdims=True)) / n1
        dx2 = (dt - t * (t * dt).sum(-1, keepdims=True)) / n2
    else:
        dx1, dx2 = da, dt
    return InfoNCEResult(float(loss), row_lse, col_lse, diag, s, dx1, dx2, dls)


def infonce_port_torch(x1, x2, logit_scale, scale_max=None, normalized=False):
    """torch CPU restatement with autograd (the reference's CPU arithmetic).

    x1, x2: torch tensors (may require grad); logit_scale: 0-d tensor (may
    require grad).  Returns the 0-d loss.  Statement order mirrors
    loss_head.py:271-283 but is written against logsumexp instead of
    nn.CrossEntropyLoss so the two can disagree only by rounding.
    """
    import torch

    if not normalized:
        x1 = x1 / x1.norm(dim=-1, keepdim=True)
        x2 = x2 / x2.norm(dim=-1, keepdim=True)
    s = logit_scale.exp().clamp(max=(scale_max or float("inf")))
    l12 = (s * x1) @ x2.t()
    l21 = (s * x2) @ x1.t()
    d = torch.diagonal(l12)
    return (torch.logsumexp(l12, dim=1) - d).mean() + (torch.logsumexp(l21, dim=1) - d).mean()


def infonce_row_sharded(x1, x2, world, logit_scale=math.log(1 / 0.07), scale_max=None,
                        normalized=False, grad_output=1.0, dtype=np.float64):
    """The multi-GPU decomposition on one process.

    Rank r owns rows [r*b, (r+1)*b) of both inputs.  After the feature
    all-gather it computes, for ITS rows only, the full row logsumexp (x1 rows
    against every x2 row) and the full column logsumexp (x2 rows against every
    x1 row); the three (B,) statistic vectors are all-gathered; each rank then
    forms dL/dx1 and dL/dx2 for its own rows from a sweep over all rows of the
    other modality.  Returns per-rank results concatenated in rank order so they
    can be compared with ``infonce_closed_form`` on the full batch.
    """
    x1 = np.asarray(x1, dtype=dtype)
    x2 = np.asarray(x2, dtype=dtype)
    B = x1.shape[0]
    assert B % world == 0
    b = B // world
    if not normalized:
        n1 = np.sqrt((x1 * x1).sum(-1, keepdims=True))
        n2 = np.sqrt((x2 * x2).sum(-1, keepdims=True))
        a_all, t_all = x1 / n1, x2 / n2          # == per-rank normalise + all-gather
    else:
        a_all, t_all = x1, x2
    s, flows = effective_scale(float(logit_scale), scale_max)
    row_lse = np.empty(B, dtype)
    col_lse = np.empty(B, dtype)
    diag = np.empty(B, dtype)
    for r in range(world):                        # forward sweeps, rank-local
        sl = slice(r * b, (r + 1) * b)
        S_rows = (dtype(s) * a_all[sl]) @ t_all.T     # (b, B)
        S_cols = (dtype(s) * t_all[sl]) @ a_all.T     # (b, B): column block, transposed
        row_lse[sl] = _lse(S_rows, 1)
        col_lse[sl] = _lse(S_cols, 1)
        diag[sl] = (S_rows[np.arange(b), np.arange(b) + r * b])
    loss = (row_lse - diag).mean() + (col_lse - diag).mean()   # after the stats all-gather
    dx1 = np.empty_like(x1)
    dx2 = np.empty_like(x2)
    dls = 0.0
    for r in range(world):                        # backward sweeps, rank-local
        sl = slice(r * b, (r + 1) * b)
        idx = np.arange(b)
        S_rows = (dtype(s) * a_all[sl]) @ t_all.T
        G = np.exp(S_rows - row_lse[sl, None]) + np.exp(S_rows - col_lse[None, :])
        G[idx, idx + r * b] -= 2.0
        G *= dtype(grad_output) / B
        da = dtype(s) * (G @ t_all)
        dls += float((G * S_rows).sum())          # scalar all-reduce across ranks
        S_cols = (dtype(s) * t_all[sl]) @ a_all.T
        Gt = np.exp(S_cols - col_lse[sl, None]) + np.exp(S_cols - row_lse[None, :])
        Gt[idx, idx + r * b] -= 2.0
        Gt *= dtype(grad_output) / B
        dt = dtype(s) * (Gt @ a_all)
        if not normalized:
            a, t = a_all[sl], t_all[sl]
            dx1[sl] = (da - a * (a * da).sum(-1, keepdims=True)) / n1[sl]
            dx2[sl] = (dt - t * (t * dt).sum(-1, keepdims=True)) / n2[sl]
        else:
            dx1[sl], dx2[sl] = da, dt
    return InfoNCEResult(float(loss), row_lse, col_lse, diag, s, dx1, dx2, dls if flows else 0.0)


def make_pair(B, D=512, rho=0.3, seed=1213, dtype="float32"):
    """Synthetic inputs of SURVEY.md section 8(d): x2 = rho*x1 + (1-rho)*randn, seed 1213."""
    import torch

    g = torch.Generator().manual_seed(seed)
    x1 = torch.randn(B, D, generator=g)
    x2 = rho * x1 + (1.0 - rho) * torch.randn(B, D, generator=g)
    return x1.numpy().astype(dtype), x2.numpy().astype(dtype)


def infonce_port_block_step(x1, x2, logit_scale, rows, scale_max=None, normalized=False):
    """One rank's share of a global-batch step in the reference's CPU arithmetic (torch fp32 + autograd).

    The reference forms the full B x B logits twice (loss_head.py:277-278); a bounded SAMPLE of that work is
    the row block `rows` (a slice): logits of those x1 rows against all x2 rows and of those x2 rows against
    all x1 rows, both cross entropies, and autograd backward -- 12*b*B*D executed flops, exactly b/B of the
    reference's 12*B^2*D.  Used only by bench.py's cpu_baseline / --impl reference legs.
    Returns the (partial) loss tensor after calling backward().
    """
    import torch

    if not normalized:
        a = x1 / x1.norm(dim=-1, keepdim=True)
        t = x2 / x2.norm(dim=-1, keepdim=True)
    else:
        a, t = x1, x2
    s = logit_scale.exp().clamp(max=(scale_max or float("inf")))
    labels = torch.arange(rows.start, rows.stop)
    l12 = (s * a[rows]) @ t.t()
    l21 = (s * t[rows]) @ a.t()
    ce = torch.nn.functional.cross_entropy
    loss = ce(l12, labels, reduction="sum") + ce(l21, labels, reduction="sum")
    loss = loss / x1.shape[0]
    loss.backward()
    return loss
