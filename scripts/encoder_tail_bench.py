"""Encoder tail (LayerNorm -> projection -> normalise -> bf16 operands) on one B200: the two fused kernels against the same span in
PyTorch eager (fp32 as the reference runs it, and bf16 autocast).  L2 flushed between timed iterations; CUDA events on the
current stream.  Writes gpurun_out/encoder_tail.json.

    python scripts/encoder_tail_bench.py [--rows 32768 4096 512] [--width 768] [--embed 512]
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as TF

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vipant_b200.encoder_tail import FusedPostEncoder  # noqa: E402


def timed(fn, flush, iters=20, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.add_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, nargs="+", default=[32768, 4096, 512])
    ap.add_argument("--width", type=int, default=768)
    ap.add_argument("--embed", type=int, default=512)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    flush = torch.zeros(96 * 1024 * 1024, device=dev)            # 384 MB > L2
    out = []
    for rows in args.rows:
        W, N = args.width, args.embed
        hidden = torch.randn(rows, W, device=dev)
        gamma = torch.rand(W, device=dev) + 0.5
        beta = torch.randn(W, device=dev) * 0.1
        proj = torch.randn(W, N, device=dev) * W ** -0.5
        post = FusedPostEncoder(None, width=W, embed_dim=N).to(dev)      # the public module (parameters require grad)
        with torch.no_grad():
            post.ln.weight.copy_(gamma)
            post.ln.bias.copy_(beta)
            post.proj.copy_(proj)

        def fused_train():                              # forward as a training step calls it (autograd state kept)
            return post(hidden)

        def fused_infer():
            with torch.no_grad():
                return post.operands(hidden, features=False)

        def eager_fp32():
            y = TF.layer_norm(hidden, (W,), gamma, beta) @ proj
            return (y / y.norm(dim=-1, keepdim=True)).to(torch.bfloat16)

        def eager_bf16():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = TF.layer_norm(hidden, (W,), gamma, beta) @ proj
            y = y.float()
            return (y / y.norm(dim=-1, keepdim=True)).to(torch.bfloat16)

        alg_train = rows * (W * 4 + W * 2 + W * 2 + N * 4 + N * 2 + 12) + W * N * 2
        alg_infer = rows * (W * 4 + W * 2 + W * 2 + N * 2 + 4) + W * N * 2
        r = {"rows": rows, "width": W, "embed": N}
        for name, fn, alg in (("fused_train", fused_train, alg_train), ("fused_infer", fused_infer, alg_infer),
                              ("eager_fp32", eager_fp32, None), ("eager_bf16_autocast", eager_bf16, None)):
            ms = timed(fn, flush)
            r[name + "_us"] = round(ms * 1e3, 2)
            if alg:
                r[name + "_alg_GBps"] = round(alg / ms / 1e6, 1)
        out.append(r)
        print(json.dumps(r), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/encoder_tail.json", "w") as f:
        json.dump({"device": torch.cuda.get_device_name(0), "note": "median of 20, L2 flushed between iterations; fused_* = FusedPostEncoder.forward (grad mode) / .operands (no_grad)", "results": out}, f,
                  indent=1)


if __name__ == "__main__":
    main()
