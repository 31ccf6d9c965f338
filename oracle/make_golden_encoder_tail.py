"""TEST INFRASTRUCTURE: golden outputs of the reference's OWN post-encoders (build container only; /root/reference is imported,
nothing is copied).

    python oracle/make_golden_encoder_tail.py      ->  tests/golden/encoder_tail.npz

``ViTPostEncoder`` (/root/reference/cvap/module/val.py:262-290) on the CLS token and ``GPTPostEncoder`` (:125-146) on the EOT token,
each followed by the heads' ``x / x.norm(dim=-1, keepdim=True)`` (clip_head.py:118), fp32 on the CPU as the reference computes it;
plus the reference autograd's gradients of sum(w * unit) for the ViT case.  Inputs are regenerated from the seed by
oracle/encoder_tail_oracle.golden_inputs (their float64 checksums are stored); outputs are stored whole (they are small).
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import reference_loader as rl  # noqa: E402
from oracle.encoder_tail_oracle import golden_inputs  # noqa: E402

CASES = {"vit": dict(seed=4101, rows=72, tokens=3, width=768, embed=512),
         "gpt": dict(seed=4102, rows=40, tokens=6, width=512, embed=512),
         "vit256": dict(seed=4103, rows=130, tokens=2, width=1024, embed=256)}
GRAD_ROWS = np.arange(1, 72, 9)


def load_reference_val():
    rl._install_stubs()
    if rl.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, rl.REFERENCE_ROOT)
    spec = importlib.util.spec_from_file_location("_vipant_reference_val", os.path.join(rl.REFERENCE_ROOT, "cvap", "module", "val.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def build(out=None):
    ref = load_reference_val()
    res = {}
    for name, c in CASES.items():
        inp = golden_inputs(**c)
        cls = ref.GPTPostEncoder if name == "gpt" else ref.ViTPostEncoder
        m = cls(None, width=c["width"], embed_dim=c["embed"])
        with torch.no_grad():
            m.ln.weight.copy_(torch.from_numpy(inp["gamma"]))
            m.ln.bias.copy_(torch.from_numpy(inp["beta"]))
            m.proj.copy_(torch.from_numpy(inp["proj"]))
        hidden = torch.from_numpy(inp["hidden"]).requires_grad_(True)
        y = m(hidden, mask=torch.from_numpy(inp["eot"])) if name == "gpt" else m(hidden)
        unit = y / y.norm(dim=-1, keepdim=True)                                   # clip_head.py:118
        res[f"{name}_y"], res[f"{name}_norm"] = y.detach().numpy(), y.detach().norm(dim=-1).numpy()    # unit rows = y / norm
        res[f"{name}_unit_rows"] = unit.detach().numpy()[::8]
        res[f"{name}_checksum"] = np.array([float(np.asarray(v, np.float64).sum()) for v in (inp["hidden"], inp["gamma"], inp["beta"], inp["proj"])])
        if name == "vit":
            (unit * torch.from_numpy(inp["w"])).sum().backward()
            res["vit_dx_rows"] = hidden.grad[:, 0, :].numpy()[GRAD_ROWS]
            res["vit_dx_norm"] = np.float64(np.linalg.norm(hidden.grad[:, 0, :].numpy().astype(np.float64)))
            res["vit_dx_other_tokens_absmax"] = np.float64(hidden.grad[:, 1:, :].abs().max())
            res["vit_dgamma"], res["vit_dbeta"] = m.ln.weight.grad.numpy(), m.ln.bias.grad.numpy()
            res["vit_dproj_rows"] = m.proj.grad.numpy()[GRAD_ROWS]
            res["vit_dproj_norm"] = np.float64(np.linalg.norm(m.proj.grad.numpy().astype(np.float64)))
    if out:
        np.savez_compressed(out, **res)
    return res


if __name__ == "__main__":
    path = os.path.join(ROOT, "tests", "golden", "encoder_tail.npz")
    build(path)
    print("wrote", path, os.path.getsize(path), "bytes")
