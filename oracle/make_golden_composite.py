"""TEST INFRASTRUCTURE: golden outputs of the reference's OWN composite heads (run unmodified through oracle/reference_loader.py).

    python oracle/make_golden_composite.py      ->  tests/golden/composite_heads.npz + composite_heads.json

VALCELossHead (va + lv + al; /root/reference/cvap/module/decoder/loss_head.py:421-495) and VACELossHead (vp / ap / va / vv / aa
with weights; :497-598) on seeded features with a DIFFERENT temperature per pair and a GradScaler-like upstream gradient:
total loss, sampled gradient rows + Frobenius norm of every input matrix, d logit_scale of every pair, and the `stats()` string.  The fused multi-pair
step of vipant_b200 (vpa_infonce_multi_fwd / _bwd) is checked against these in tests/test_gpu_composite.py.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.reference_loader import Cfg, load_reference_loss_head  # noqa: E402

B, D, SEED = 320, 512, 1213
VAL_SCALES = {"va": 2.0, "lv": 2.659260036932778, "al": 3.2}            # ln of the temperatures (one per pair)
VA_SCALES = {"vp": 2.1, "ap": 2.659260036932778, "va": 2.9, "vv": 2.4, "aa": 3.5}
VA_WEIGHTS = dict(vp_w=1.0, ap_w=0.5, va_w=2.0, vv_w=0.25, aa_w=0.75)
GRAD_OUT = 128.0
ROWS = np.arange(3, B, 7)                   # sampled gradient rows kept in the fixture (+ the Frobenius norm of every matrix)


def features(n):
    g = torch.Generator().manual_seed(SEED)
    base = torch.randn(B, D, generator=g)
    return [(0.4 * base + torch.randn(B, D, generator=g)) for _ in range(n)]          # correlated rows: a non-trivial diagonal


def run(head, scales, feats):
    for key, v in scales.items():
        with torch.no_grad():
            getattr(head, "loss_head_" + key).logit_scale.fill_(v)
    xs = [f.clone().requires_grad_(True) for f in feats]
    loss = head(*xs, normalized=False, names=None)
    (loss * GRAD_OUT).backward()
    return dict(loss=float(loss), stats=head.stats(nstep=1), dx=[x.grad.numpy() for x in xs],
                dls={k: float(getattr(head, "loss_head_" + k).logit_scale.grad) for k in scales})


def main():
    ref = load_reference_loss_head()
    out, meta = {}, dict(B=B, D=D, seed=SEED, grad_out=GRAD_OUT, val_scales=VAL_SCALES, va_scales=VA_SCALES, va_weights=VA_WEIGHTS)
    head = ref.VALCELossHead(Cfg(scaling=True, scale_max=None, va=True, lv=True, al=True)).train()
    r = run(head, VAL_SCALES, features(3))
    meta["val"] = dict(loss=r["loss"], stats=r["stats"], dls=r["dls"])
    for i, g in enumerate(r["dx"]):
        out[f"val_dx{i}"], out[f"val_norm{i}"] = g[ROWS], np.float64(np.linalg.norm(g.astype(np.float64)))
    head = ref.VACELossHead(Cfg(scaling=True, scale_max=None, vp=True, ap=True, va=True, vv=True, aa=True, **VA_WEIGHTS)).train()
    r = run(head, VA_SCALES, features(5))
    meta["va"] = dict(loss=r["loss"], stats=r["stats"], dls=r["dls"])
    for i, g in enumerate(r["dx"]):
        out[f"va_dx{i}"], out[f"va_norm{i}"] = g[ROWS], np.float64(np.linalg.norm(g.astype(np.float64)))
    out["rows"] = ROWS
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "composite_heads.npz"), **out)
    with open(os.path.join(ROOT, "tests", "golden", "composite_heads.json"), "w") as fw:
        json.dump(meta, fw, indent=1)
    print(json.dumps({k: meta[k] for k in ("val", "va")}, indent=1))


if __name__ == "__main__":
    main()
