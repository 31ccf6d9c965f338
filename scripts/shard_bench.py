"""One rank's share of an R-way row-sharded step on ONE GPU (no collectives): b = B/R local rows against B global rows.
Used to tune the per-rank kernel sequence and host overhead of the multi-GPU path without holding R GPUs.
    python scripts/shard_bench.py [R] [B] [steps]"""
import ctypes, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vipant_b200 import _cabi, functional as F_

R = int(sys.argv[1]) if len(sys.argv) > 1 else 8
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 30
D, b = 512, B // R
dev = torch.device("cuda", 0)
g = torch.Generator(device="cuda").manual_seed(0)
x1 = torch.randn(b, D, device=dev, generator=g); x2 = 0.3 * x1 + 0.7 * torch.randn(b, D, device=dev, generator=g)
full = torch.randn(2, B, D, device=dev, generator=g)
full = (full / full.norm(dim=-1, keepdim=True)).to(torch.bfloat16)
ls = torch.tensor(math.log(1 / 0.07), device=dev)
K = F_._KERNELS
prec = _cabi.PREC_BF16_TC
lib = _cabi.lib()
gout = torch.tensor(1.0, device=dev)

def step():
    a, t, inv, dcos = K.normalize_pair(x1, x2, False, prec)
    a_all, t_all = full[0], full[1]
    a_all[:b].copy_(a); t_all[:b].copy_(t)          # stands in for the all-gather
    col_sum, ws = K.forward_sweep(a, t, a_all, t_all, 0, ls, None, prec)
    stats, scale = K.forward_finish(b, B, D, 0, ls, None, dcos, prec, ws, col_sum)
    stats_all = torch.empty(3, B, device=dev); stats_all[:, :b] = stats; stats_all[:, b:] = stats[:, :1].mean()
    loss = K.loss(stats_all)
    return K.backward(x1, x2, a, t, a_all, t_all, inv, stats_all, scale, ws, 0, gout, False, prec)

for _ in range(5): step()
torch.cuda.synchronize()
lib.vpa_profile_enable(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(steps): step()
e1.record(); t_host = time.perf_counter() - t0
torch.cuda.synchronize()
out = {}
for kind, name in ((0, "normalize"), (1, "fwd_sweep"), (2, "bwd_sweep"), (5, "fwd_gated"), (6, "finalize")):
    tot, n = ctypes.c_float(), ctypes.c_int()
    lib.vpa_profile_read(kind, ctypes.byref(tot), ctypes.byref(n))
    out[name] = round(tot.value / max(n.value, 1), 4)
print(f"R={R} b={b} B={B}: gpu {e0.elapsed_time(e1)/steps:.3f} ms/step, host enqueue {t_host/steps*1e3:.3f} ms/step, kernels {out} sum {sum(out.values()):.3f}")
