"""TEST INFRASTRUCTURE -- CPU restatement (numpy, float64) of the reference's encoder tail, the span `vpa_encoder_tail` replaces.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this; the product path never does.

Follows, line by line:
  * ``ViTPostEncoder.forward``  /root/reference/cvap/module/val.py:288-290   ``x = self.ln(x[:, 0, :]); x = x @ self.proj``
  * ``GPTPostEncoder.forward``  /root/reference/cvap/module/val.py:143-146   ``x = self.ln(x); x = x[arange, mask] @ self.proj``
  * ``LayerNorm``               /root/reference/clip/model.py:154-160        torch.nn.LayerNorm computed in fp32 (biased variance,
                                                                             eps inside the square root, elementwise affine)
  * the heads' normalisation    /root/reference/cvap/module/encoder/clip_head.py:117-118, audio_head.py:209-210
                                ``x = x / x.norm(dim=-1, keepdim=True)``
Pinned against the reference's own modules executed here: oracle/make_golden_encoder_tail.py -> tests/golden/encoder_tail.npz.
"""
from __future__ import annotations

import numpy as np

EPS = 1e-5                      # torch.nn.LayerNorm default, which the reference's LayerNorm(width) keeps


def layer_norm(x, gamma, beta, eps=EPS):
    x = np.asarray(x, np.float64)
    mean = x.mean(-1, keepdims=True)
    var = ((x - mean) ** 2).mean(-1, keepdims=True)           # biased
    return (x - mean) / np.sqrt(var + eps) * np.asarray(gamma, np.float64) + np.asarray(beta, np.float64)


def select_rows(hidden, mask=None):
    """CLS token (val.py:288) or, with the EOT positions ``mask``, ``x[arange, mask]`` (val.py:145)."""
    hidden = np.asarray(hidden)
    if hidden.ndim == 2:
        return hidden
    if mask is None:
        return hidden[:, 0, :]
    return hidden[np.arange(hidden.shape[0]), np.asarray(mask)]


def encoder_tail(hidden, gamma, beta, proj, mask=None, eps=EPS):
    """-> (y, unit): projected features (val.py:289 / :145) and their unit rows (clip_head.py:118), float64.
    LayerNorm acts per token, so selecting the row before or after it (ViT vs GPT order) is the same arithmetic."""
    ln = layer_norm(select_rows(hidden, mask), gamma, beta, eps)
    y = ln @ np.asarray(proj, np.float64)
    return y, y / np.linalg.norm(y, axis=-1, keepdims=True)


def encoder_tail_grads(hidden2d, gamma, beta, proj, w, eps=EPS):
    """Gradients of  sum(w * unit)  w.r.t. (x, gamma, beta, proj), float64 -- the closed forms the backward uses."""
    x = np.asarray(hidden2d, np.float64)
    gamma = np.asarray(gamma, np.float64)
    proj = np.asarray(proj, np.float64)
    w = np.asarray(w, np.float64)
    mean = x.mean(-1, keepdims=True)
    rstd = 1.0 / np.sqrt(((x - mean) ** 2).mean(-1, keepdims=True) + eps)
    xhat = (x - mean) * rstd
    ln = xhat * gamma + np.asarray(beta, np.float64)
    y = ln @ proj
    nrm = np.linalg.norm(y, axis=-1, keepdims=True)
    u = y / nrm
    dy = (w - u * (w * u).sum(-1, keepdims=True)) / nrm        # Jacobian of y / ||y||
    dproj = ln.T @ dy
    dln = dy @ proj.T
    dgamma = (dln * xhat).sum(0)
    dbeta = dln.sum(0)
    dxh = dln * gamma
    dx = rstd * (dxh - dxh.mean(-1, keepdims=True) - xhat * (dxh * xhat).mean(-1, keepdims=True))
    return dx, dgamma, dbeta, dproj


def golden_inputs(seed, rows, tokens, width, embed):
    """Seeded inputs shared by the generator and the tests (numpy's PCG64 stream is platform independent), float32."""
    rng = np.random.default_rng(seed)
    hidden = (rng.standard_normal((rows, tokens, width)) * 1.5 + 0.25).astype(np.float32)
    gamma = (1.0 + 0.2 * rng.standard_normal(width)).astype(np.float32)
    beta = (0.1 * rng.standard_normal(width)).astype(np.float32)
    proj = (width ** -0.5 * rng.standard_normal((width, embed))).astype(np.float32)
    eot = rng.integers(0, tokens, size=rows).astype(np.int64)
    w = rng.standard_normal((rows, embed)).astype(np.float32)
    return dict(hidden=hidden, gamma=gamma, beta=beta, proj=proj, eot=eot, w=w)
