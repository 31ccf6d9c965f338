"""Developer diagnostics on a B200 (not a test): each stage compares one kernel with the CPU oracle.

    python scripts/gpu_check.py <stage> [B] [D]
stages: normalize | simt | tc_fwd | tc_bwd | retrieval | host
Run every stage in its own process (a device-side trap poisons the CUDA context).
"""
import math
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vipant_b200 as vb  # noqa: E402
from vipant_b200 import _cabi, functional as F_  # noqa: E402
from oracle import infonce_oracle as io, retrieval_oracle as ro  # noqa: E402


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def run_infonce(precision, B, D, ls=math.log(1 / 0.07), rho=0.3, normalized=False, gout=1.0, scale_max=None):
    x1n, x2n = io.make_pair(B, D, rho, 1213)
    if normalized:
        x1n /= np.linalg.norm(x1n, axis=-1, keepdims=True); x2n /= np.linalg.norm(x2n, axis=-1, keepdims=True)
    ref = io.infonce_closed_form(x1n, x2n, ls, scale_max, normalized, gout)
    x1 = torch.from_numpy(x1n).cuda().requires_grad_(True)
    x2 = torch.from_numpy(x2n).cuda().requires_grad_(True)
    lsc = torch.tensor(ls, device="cuda", requires_grad=True)
    t0 = time.time()
    loss = vb.infonce_loss(x1, x2, lsc, scale_max=scale_max, normalized=normalized, precision=precision)
    (loss * gout).backward()
    torch.cuda.synchronize()
    dt = time.time() - t0
    print(f"[{precision}] B={B} D={D} ls={ls:.3f} norm={normalized} g={gout}: loss {loss.item():.6f} vs {ref.loss:.6f} "
          f"(rel {abs(loss.item() - ref.loss) / abs(ref.loss):.2e})  dls {lsc.grad.item():.6g} vs {ref.dlogit_scale:.6g}  "
          f"dx1 rel {rel(x1.grad.cpu().numpy(), ref.dx1):.2e} dx2 rel {rel(x2.grad.cpu().numpy(), ref.dx2):.2e}  [{dt*1e3:.1f} ms]",
          flush=True)


def main():
    stage = sys.argv[1]
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    D = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    print("device:", torch.cuda.get_device_name(0), "lib version", _cabi.lib().vpa_version(), flush=True)
    if stage == "normalize":
        for dt in (torch.float32, torch.bfloat16, torch.float16):
            x = torch.randn(1000, D, device="cuda").to(dt)
            y = vb.l2_normalize(x)
            r = x.float() / x.float().norm(dim=-1, keepdim=True)
            print(dt, "max abs err", (y - r).abs().max().item(), flush=True)
    elif stage == "simt":
        run_infonce("fp32", 64, 512)
        run_infonce("fp32", B, D)
        run_infonce("fp32", 200, 256, ls=math.log(100.0) + 0.1, scale_max=100.0, normalized=True, rho=0.06)
        run_infonce("fp32", 1000, 128, gout=65536.0, rho=0.0)
    elif stage == "tc_fwd":
        x1n, x2n = io.make_pair(B, D, 0.3, 1213)
        ref = io.infonce_closed_form(x1n, x2n)
        x1 = torch.from_numpy(x1n).cuda(); x2 = torch.from_numpy(x2n).cuda()
        with torch.no_grad():
            loss = vb.infonce_loss(x1, x2, torch.tensor(math.log(1 / 0.07), device="cuda"), precision="bf16")
        torch.cuda.synchronize()
        print(f"tc fwd B={B} D={D}: loss {loss.item():.6f} vs {ref.loss:.6f} rel {abs(loss.item()-ref.loss)/abs(ref.loss):.2e}", flush=True)
    elif stage == "tc_bwd":
        run_infonce("bf16", B, D)
        run_infonce("bf16", 200, 256, ls=math.log(100.0) + 0.1, scale_max=100.0, normalized=True, rho=0.06)
        run_infonce("bf16", 1000, 128, gout=65536.0, rho=0.0)
        run_infonce("bf16", 2048, 512)
    elif stage == "retrieval":
        from oracle.make_golden import retrieval_inputs_1v5
        a, t = retrieval_inputs_1v5(n=150, seed=1216)
        head = vb.CELossHead(type("C", (), dict(scaling=True, scale_max=None))()).cuda().eval()
        for i in range(0, 150, 64):
            head(torch.from_numpy(a[i:i + 64]).cuda(), torch.from_numpy(t[i * 5:(i + 64) * 5]).cuda())
        rep = head.report()
        exp, r12, r21 = ro.report(ro.normalize(a), ro.normalize(t))
        print(rep); print("match oracle string:", rep == exp, flush=True)
    elif stage == "host":
        lib = _cabi.lib()
        x1n, x2n = io.make_pair(B, D, 0.3, 1213)
        ref = io.infonce_closed_form(x1n, x2n)
        import ctypes
        for prec in (0, 1):
            nbytes = lib.vpa_infonce_host_scratch_bytes(B, D, prec)
            scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
            loss = ctypes.c_float(); dls = ctypes.c_float()
            dx1 = np.empty_like(x1n); dx2 = np.empty_like(x2n)
            rc = lib.vpa_infonce_step_host(x1n.ctypes.data, x2n.ctypes.data, B, D, math.log(1 / 0.07), 0.0, 1.0, prec,
                                           scratch.data_ptr(), nbytes, ctypes.addressof(loss), ctypes.addressof(dls),
                                           dx1.ctypes.data, dx2.ctypes.data, None)
            print(f"host step prec={prec} rc={rc} loss {loss.value:.6f} vs {ref.loss:.6f} dls {dls.value:.6g} vs {ref.dlogit_scale:.6g} "
                  f"dx1 rel {rel(dx1, ref.dx1):.2e}", flush=True)


if __name__ == "__main__":
    main()
