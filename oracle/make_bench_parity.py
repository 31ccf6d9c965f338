"""TEST INFRASTRUCTURE (checker, never shipped): fp64 ground truth of the HEADLINE benchmark step.

    python oracle/make_bench_parity.py            ->  tests/golden/bench_parity_b32768.npz

bench.py prints a `parity` object at every N (1, 2, 4, 8 GPUs): loss, d logit_scale, the Frobenius norm of every 4096-row
block of dx1 / dx2 and 32 sampled gradient rows per block, compared with the values committed here.  They are the SURVEY.md
section 8 closed form of the reference's loss (/root/reference/cvap/module/decoder/loss_head.py:271-283 + autograd, checked
against the reference's own autograd in tests/test_oracle.py) evaluated in float64 on the benchmark's inputs
(bench.make_inputs: seed 1213, x2 = 0.3 x1 + 0.7 randn, B = 32768, D = 512, logit_scale = log(1/0.07), grad_output = 1),
blockwise so that the 32768 x 32768 logits never exist at once.  Takes a few minutes on 8 cores.
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

BLOCK = 4096            # rows per rank at 8 GPUs: the granularity of the per-block norms
SAMPLES = 32            # sampled rows per block


def sample_rows(B):
    """Fixed sample: SAMPLES rows of every BLOCK-row block (stride chosen odd so that rows hit all tile positions)."""
    rows = []
    for blk in range(B // BLOCK):
        rows += [blk * BLOCK + (7 + 127 * i) % BLOCK for i in range(SAMPLES)]
    return np.array(sorted(rows), dtype=np.int64)


def blockwise_truth(x1, x2, rb=2048):
    """(loss, d logit_scale, dx1, dx2) in float64, `rb` rows of the logits at a time (torch tensors in, torch tensors out)."""
    B, D = x1.shape
    x1, x2 = x1.double(), x2.double()
    n1, n2 = x1.norm(dim=-1, keepdim=True), x2.norm(dim=-1, keepdim=True)
    a, t = x1 / n1, x2 / n2
    s = float(np.exp(np.float32(math.log(1 / 0.07))))          # the fp32 parameter value the kernels read
    row_lse = torch.empty(B, dtype=torch.float64)
    cmax = torch.full((B,), -float("inf"), dtype=torch.float64)
    csum = torch.zeros(B, dtype=torch.float64)
    diag = s * (a * t).sum(-1)
    for r0 in range(0, B, rb):                                   # pass 1: statistics
        S = s * (a[r0:r0 + rb] @ t.T)
        row_lse[r0:r0 + rb] = torch.logsumexp(S, dim=1)
        m = S.max(dim=0).values
        new = torch.maximum(cmax, m)
        csum = csum * torch.exp(cmax - new) + torch.exp(S - new[None, :]).sum(0)
        cmax = new
        if B >= 8192:
            print(f"pass 1 rows {r0}", flush=True)
    col_lse = cmax + torch.log(csum)
    loss = float(((row_lse - diag) + (col_lse - diag)).mean())
    da = torch.empty(B, D, dtype=torch.float64)
    dt = torch.zeros(B, D, dtype=torch.float64)
    dls = 0.0
    for r0 in range(0, B, rb):                                   # pass 2: G = (softmax_rows + softmax_cols - 2 I) / B
        S = s * (a[r0:r0 + rb] @ t.T)
        G = torch.exp(S - row_lse[r0:r0 + rb, None]) + torch.exp(S - col_lse[None, :])
        idx = torch.arange(min(rb, B - r0))
        G[idx, idx + r0] -= 2.0
        G /= B
        dls += float((G * S).sum())
        da[r0:r0 + rb] = s * (G @ t)
        dt += s * (G.T @ a[r0:r0 + rb])
        if B >= 8192:
            print(f"pass 2 rows {r0}", flush=True)
    dx1 = (da - a * (a * da).sum(-1, keepdim=True)) / n1
    dx2 = (dt - t * (t * dt).sum(-1, keepdim=True)) / n2
    return loss, dls, dx1, dx2


def main(B=32768, D=512, rb=2048):
    import bench
    x1, x2 = bench.make_inputs(B, D, 0, B)
    loss, dls, dx1, dx2 = blockwise_truth(x1, x2, rb)
    rows = sample_rows(B)
    nb = B // BLOCK
    out = dict(
        B=B, D=D, block=BLOCK, loss=loss, dlogit_scale=dls, rows=rows,
        dx1_rows=dx1[rows].numpy().astype(np.float32), dx2_rows=dx2[rows].numpy().astype(np.float32),
        dx1_block_norm=np.array([float(dx1[k * BLOCK:(k + 1) * BLOCK].norm()) for k in range(nb)]),
        dx2_block_norm=np.array([float(dx2[k * BLOCK:(k + 1) * BLOCK].norm()) for k in range(nb)]))
    path = os.path.join(ROOT, "tests", "golden", f"bench_parity_b{B}.npz")
    np.savez_compressed(path, **out)
    print(path, "loss", loss, "dlogit_scale", dls)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    main(*(int(v) for v in sys.argv[1:3]))
