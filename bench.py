#!/usr/bin/env python
"""Headline benchmark: fused InfoNCE forward+backward, global batch 32768 x 512, bf16 tensor-core mode.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--dim D]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one synthetic batch:
  normalise+cast both modalities -> [operand exchange] -> single-pass forward statistics -> loss -> backward (dx1, dx2,
  dlogit_scale).
metric  pairs/s = GLOBAL batch / step time   (strong scaling: the global batch is fixed at 32768, rows are
        sharded over the N ranks, BASELINE.json configs[2]); at N > 1 the exchange runs over NVLink peer memory
        (config.transport = "p2p"; VIPANT_TRANSPORT=nccl|host selects the other transports)
value   inputs already resident in HBM, CUDA-event timed, max over ranks
e2e     the same step through the host-buffer entry (C-ABI vpa_infonce_step_host at N=1 -- pipelined over row shards on
        three streams --, the Python public API with pinned host tensors at N>1): H2D of the embeddings and D2H of
        loss + gradients inside the timed region
roofline  the backward sweep kernel (tcgen05), algorithmic 6*b*B*D flops per launch / its CUDA-event time, against the
        measured cuBLAS bf16 peak: the BURST figure unless this run's own clock samples show a power-capped median (then
        the sustained one); both fractions are printed.  traffic = DRAM bytes per launch from the committed ncu capture
parity  every rank compares its loss, d logit_scale and gradients (Frobenius norm of every 4096-row block + 32 sampled rows
        per block) with the fp64 ground truth committed in tests/golden/bench_parity_b32768.npz (oracle/make_bench_parity.py);
        the run FAILS (exit code 3) outside the bf16-mode bars (loss 1e-3, gradients 1e-2)
eager_b200  the reference's formula (loss_head.py:271-283) in PyTorch eager on the same B200: fp32 and bf16 autocast
cpu_baseline  the oracle's torch-CPU port of the reference arithmetic on a bounded, FIXED row-block sample
c2_batch512_latency  BASELINE.json configs[1] (per-GPU batch 512 x 512): microseconds per step, eager and CUDA-graph replay
--impl reference   times that CPU port as the main line (the reference itself cannot travel to the GPU box)
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "InfoNCE fwd+bwd pairs/sec at 32Kx512"
UNIT = "pairs/s"
SEED = 1213            # the reference's default seed (configs/default.yaml:9)
RHO = 0.3
PROF_EVERY = 4         # steps of the timed region that carry the per-kernel CUDA events


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32768, help="GLOBAL batch (rows of each modality)")
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-eager", action="store_true")
    return ap.parse_args()


def peaks(clocks=None):
    """Measured peaks (driver-written MEASURED_PEAKS.json).  bf16: the burst figure is the roofline of a kernel that runs at
    the boost clock; only when this run's own clock samples show a power-capped median (< 85 % of max) does the sustained
    figure (measured at a ~1.34 GHz median) apply."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fr:
            p = json.load(fr)
        burst = float(p.get("bf16_tflops", 1660.0))
        sustained = float(p.get("bf16_tflops_sustained", burst))
        capped = bool(clocks and clocks.get("sm_mhz") and clocks.get("sm_max_mhz")
                      and clocks["sm_mhz"] < 0.85 * clocks["sm_max_mhz"])
        return dict(tflops=sustained if capped else burst, burst=burst, sustained=sustained,
                    hbm=float(p.get("hbm_gbs", 6650.0)),
                    source="measured (MEASURED_PEAKS.json, " + ("sustained bf16: power-capped clocks in this run)" if capped
                                                               else "burst bf16: boost clocks in this run)"))
    return dict(tflops=1590.0, burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


def make_inputs(B, D, lo, hi):
    """Rows [lo, hi) of the seeded global batch: x2 = rho*x1 + (1-rho)*randn."""
    g = torch.Generator().manual_seed(SEED)
    x1 = torch.randn(B, D, generator=g)
    x2 = RHO * x1 + (1.0 - RHO) * torch.randn(B, D, generator=g)
    return x1[lo:hi].contiguous(), x2[lo:hi].contiguous()


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)          # warm the query path
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.002)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------- CPU arm
CPU_SAMPLE_ROWS = 2048      # fixed: 12 * 2048 * B * D executed flops per step (1/16 of the reference's full step at B = 32768)


def cpu_block_sample(B, D):
    """Reference arithmetic (oracle torch-CPU port) on a FIXED row block of the SAME global batch, all host threads."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)          # (torch.distributed.run exports OMP_NUM_THREADS=1; this call overrides it)
    x1, x2 = make_inputs(B, D, 0, B)
    x1.requires_grad_(True)
    x2.requires_grad_(True)
    ls = torch.tensor(math.log(1 / 0.07), requires_grad=True)
    return x1, x2, ls, min(B, CPU_SAMPLE_ROWS), torch.get_num_threads()


def run_cpu_steps(x1, x2, ls, b, steps):
    from oracle import infonce_oracle as io
    times = []
    for _ in range(steps):
        x1.grad = x2.grad = ls.grad = None
        t0 = time.perf_counter()
        io.infonce_port_block_step(x1, x2, ls, slice(0, b))
        times.append(time.perf_counter() - t0)
    return times


def reference_arm(args, rank, world):
    if rank != 0:
        return
    B, D = args.batch, args.dim
    x1, x2, ls, b, cores = cpu_block_sample(B, D)
    run_cpu_steps(x1, x2, ls, b, max(1, min(args.warmup, 1)))
    steps = max(1, args.steps)
    t0 = time.perf_counter()
    times = []
    for _ in range(steps):
        times += run_cpu_steps(x1, x2, ls, b, 1)
        if time.perf_counter() - t0 > 150:           # keep the whole run within a few minutes
            break
    ms = float(np.mean(times) * 1e3)
    value = b / (ms * 1e-3)
    sample = (f"fixed row block of {b} of the {B} global rows per step (both logits blocks {b}x{B}, 2x cross entropy, autograd "
              f"backward; 12*b*B*D executed flops = b/B of the reference's full step), fp32 torch CPU, {cores} torch threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": 1, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"InfoNCE fwd+bwd, global batch {B} x {D}, fp32 CPU reference arithmetic", "global_batch": B,
                   "dim": D, "sample_rows": b},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count(), "torch_threads": cores},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- small-batch latency (configs[1])
def c2_latency(lib, dev, iters=300):
    """BASELINE.json configs[1]: per-GPU batch 512 x 512, bf16 mode.  1 GFLOP -- launch-latency bound, so it is reported in
    microseconds per fwd+bwd step, eager (two C-ABI calls = 9 kernels) and replayed from a CUDA graph of the same calls."""
    from vipant_b200 import _cabi
    B, D, prec = 512, 512, _cabi.PREC_BF16_TC
    x1h, x2h = make_inputs(B, D, 0, B)
    x1, x2 = x1h.to(dev), x2h.to(dev)
    ls = torch.tensor(math.log(1 / 0.07), device=dev)
    g = torch.tensor(1.0, device=dev)
    nbytes = lib.vpa_sharded_state_bytes(B, 1, D, prec)
    state = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    loss, dls = torch.empty((), device=dev), torch.empty((), device=dev)
    dx1, dx2 = torch.empty_like(x1), torch.empty_like(x2)

    def enqueue():
        st = torch.cuda.current_stream().cuda_stream
        _cabi.check(lib.vpa_infonce_fwd_sharded(None, x1.data_ptr(), x2.data_ptr(), _cabi.F32, B, 1, 0, D, D, D, 0, ls.data_ptr(),
                                                0.0, prec, state.data_ptr(), nbytes, loss.data_ptr(), st), "fwd")
        _cabi.check(lib.vpa_infonce_bwd_sharded(None, x1.data_ptr(), x2.data_ptr(), _cabi.F32, B, 1, 0, D, D, D, 0, prec,
                                                g.data_ptr(), state.data_ptr(), nbytes, dx1.data_ptr(), dx2.data_ptr(),
                                                dls.data_ptr(), 1, st), "bwd")

    def timed(fn):
        for _ in range(20):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e3

    out = {"batch": B, "dim": D, "eager_us_per_step": timed(enqueue), "loss": float(loss.item())}
    try:
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            enqueue()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            enqueue()
        out["cuda_graph_us_per_step"] = timed(graph.replay)
        out["graph_loss"] = float(loss.item())
    except Exception as exc:          # reported, never fatal for the headline line
        out["cuda_graph_us_per_step"] = None
        out["cuda_graph_error"] = str(exc)[:200]
    return out


# ------------------------------------------------------------------------------------------- parity (driver-visible, every N)
def parity_check(B, D, rank, b, loss_val, dls_val, dx1, dx2, dev, world):
    """Compare this rank's results with the committed fp64 ground truth; all-reduce the worst case over the ranks."""
    path = os.path.join(ROOT, "tests", "golden", f"bench_parity_b{B}.npz")
    if D != 512 or not os.path.exists(path):
        return None
    ref = np.load(path)
    blk = int(ref["block"])
    lo, hi = rank * b, (rank + 1) * b
    worst_rows = worst_norm = 0.0
    rows = ref["rows"]
    sel = (rows >= lo) & (rows < hi)
    for dx, key in ((dx1, "dx1"), (dx2, "dx2")):
        if sel.any():
            got = dx[torch.from_numpy(rows[sel] - lo).to(dev)].double().cpu().numpy()
            want = ref[key + "_rows"][sel].astype(np.float64)
            worst_rows = max(worst_rows, float(np.linalg.norm(got - want) / np.linalg.norm(want)))
        for k in range(lo // blk, hi // blk):       # (b is a multiple of the block for N <= 8)
            n = float(dx[k * blk - lo:(k + 1) * blk - lo].double().norm().item())
            want = float(ref[key + "_block_norm"][k])
            worst_norm = max(worst_norm, abs(n - want) / want)
    loss_rel = abs(loss_val - float(ref["loss"])) / abs(float(ref["loss"]))
    dls_rel = abs(dls_val - float(ref["dlogit_scale"])) / abs(float(ref["dlogit_scale"]))
    v = torch.tensor([loss_rel, dls_rel, worst_rows, worst_norm], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
    loss_rel, dls_rel, worst_rows, worst_norm = (float(x) for x in v.tolist())
    ok = loss_rel <= 1e-3 and dls_rel <= 1e-2 and worst_rows <= 1e-2 and worst_norm <= 1e-2
    return {"loss_rel": loss_rel, "dls_rel": dls_rel, "grad_rel": worst_rows, "grad_block_norm_rel": worst_norm, "ok": bool(ok),
            "bars": {"loss": 1e-3, "grad": 1e-2}, "rows_per_rank_checked": int(sel.sum()), "ranks": world,
            "reference": "fp64 closed form of loss_head.py:271-283 on the same inputs (tests/golden/bench_parity_b%d.npz, "
                         "oracle/make_bench_parity.py)" % B}


# ------------------------------------------------------------------------------------------- same-box PyTorch eager baseline
def eager_b200(dev, D, sizes=(32768, 512), steps=5):
    """The reference's formula (loss_head.py:271-283: normalise, exp, BOTH logits matrices, 2 x CrossEntropyLoss, autograd)
    executed by PyTorch eager on this B200 -- fp32 (TF32 off: torch's default for matmul) and under bf16 autocast.  This is
    the same-hardware baseline; it materialises the B x B logits (2 x 4 GiB fp32 at 32768 plus softmax buffers)."""
    out = {}
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    ce = torch.nn.CrossEntropyLoss()
    try:
        for B in sizes:
            h1, h2 = make_inputs(B, D, 0, B)
            x1, x2 = h1.to(dev).requires_grad_(True), h2.to(dev).requires_grad_(True)
            ls = torch.tensor(math.log(1 / 0.07), device=dev, requires_grad=True)
            labels = torch.arange(B, device=dev)

            def step():
                x1.grad = x2.grad = ls.grad = None
                a = x1 / x1.norm(dim=-1, keepdim=True)
                t = x2 / x2.norm(dim=-1, keepdim=True)
                sc = ls.exp()
                loss = ce(sc * a @ t.t(), labels) + ce(sc * t @ a.t(), labels)
                loss.backward()
                return loss

            def amp():
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return step()

            rec = {}
            for name, fn in (("fp32_ms", step), ("bf16_autocast_ms", amp)):
                try:
                    n = steps if B > 4096 else 100
                    for _ in range(3):
                        fn()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    torch.cuda.synchronize()
                    e0.record()
                    for _ in range(n):
                        loss = fn()
                    e1.record()
                    torch.cuda.synchronize()
                    rec[name] = e0.elapsed_time(e1) / n
                    rec[name.replace("_ms", "_loss")] = float(loss)
                except Exception as exc:      # e.g. out of memory for the materialised logits
                    rec[name] = None
                    rec[name.replace("_ms", "_error")] = str(exc)[:160]
            out[f"B{B}"] = rec
            del x1, x2
            torch.cuda.empty_cache()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32
    return out


# ------------------------------------------------------------------------------------------- GPU arm
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (vipant_b200 has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import vipant_b200 as vb
    from vipant_b200 import _cabi
    lib = _cabi.lib()

    from vipant_b200 import functional as _vf
    transport = _vf._transport(dist.group.WORLD if world > 1 else None)
    B, D = args.batch, args.dim
    assert B % world == 0, "global batch must divide by the number of ranks"
    b = B // world
    x1h, x2h = make_inputs(B, D, rank * b, (rank + 1) * b)
    x1 = x1h.to(dev).requires_grad_(True)
    x2 = x2h.to(dev).requires_grad_(True)
    ls = torch.tensor(math.log(1 / 0.07), device=dev, requires_grad=True)
    gout = torch.tensor(1.0, device=dev)
    group = dist.group.WORLD if world > 1 else None

    def step():
        x1.grad = x2.grad = ls.grad = None
        loss = vb.infonce_loss(x1, x2, ls, precision=args.precision, group=group)
        loss.backward(gout)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        loss = step()
    barrier()

    launches0 = lib.vpa_launch_count()
    sampler = ClockSampler(local_rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lib.vpa_profile_enable(1)
    barrier()
    sampler.start()
    th0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        # kernel attribution: the library brackets its dominant kernels with CUDA events on the launch stream on every
        # PROF_EVERY-th step of the timed region (an event record is a stream operation of its own: ~10 of them per step
        # are a measurable share of a 0.6 ms step at 8 GPUs, so the other steps run uninstrumented)
        lib.vpa_profile_hold(0 if i % PROF_EVERY == 0 else 1)
        loss = step()
    lib.vpa_profile_hold(0)
    e1.record()
    host_ms = (time.perf_counter() - th0) / args.steps * 1e3       # host enqueue time per step (no sync inside)
    launches = int(lib.vpa_launch_count() - launches0)             # counted by the library at every launch site
    barrier()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    prof = {}
    for kind, name in ((0, "normalize"), (1, "fwd_sweep"), (2, "bwd_sweep"), (5, "fwd_general_gated"), (6, "finalize"), (7, "push")):
        tot, n = ctypes.c_float(), ctypes.c_int()
        lib.vpa_profile_read(kind, ctypes.byref(tot), ctypes.byref(n))
        prof[name] = (tot.value, n.value)
    lib.vpa_profile_enable(0)
    loss_val = float(loss.item())
    # ---- parity of the timed step's results against the committed fp64 ground truth (outside the timed region)
    parity = parity_check(B, D, rank, b, loss_val, float(ls.grad.item()), x1.grad.detach(), x2.grad.detach(), dev, world)

    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    value = B / (ms_step * 1e-3)

    # ---- e2e: host buffers in, loss + gradients out, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(3, min(args.steps, 10))
        h2d = 2 * b * D * 4
        d2h = 2 * b * D * 4 + 8
        if world == 1:
            prec = _cabi.PREC_BF16_TC if args.precision == "bf16" else _cabi.PREC_FP32_SIMT
            nbytes = lib.vpa_infonce_host_scratch_bytes(B, D, prec)
            scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            hx1, hx2 = x1h.pin_memory(), x2h.pin_memory()
            hd1, hd2 = torch.empty_like(hx1).pin_memory(), torch.empty_like(hx2).pin_memory()
            hl, hs = ctypes.c_float(), ctypes.c_float()

            def e2e_step():
                rc = lib.vpa_infonce_step_host(hx1.data_ptr(), hx2.data_ptr(), B, D, math.log(1 / 0.07), 0.0, 1.0, prec,
                                               scratch.data_ptr(), nbytes, ctypes.addressof(hl), ctypes.addressof(hs),
                                               hd1.data_ptr(), hd2.data_ptr(), torch.cuda.current_stream().cuda_stream)
                _cabi.check(rc, "vpa_infonce_step_host")
                return hl.value
        else:
            hx1, hx2 = x1h.pin_memory(), x2h.pin_memory()
            hd1, hd2 = torch.empty_like(hx1).pin_memory(), torch.empty_like(hx2).pin_memory()

            def e2e_step():
                d1 = hx1.to(dev, non_blocking=True).requires_grad_(True)
                d2 = hx2.to(dev, non_blocking=True).requires_grad_(True)
                ls.grad = None
                l = vb.infonce_loss(d1, d2, ls, precision=args.precision, group=group)
                l.backward(gout)
                hd1.copy_(d1.grad, non_blocking=True)
                hd2.copy_(d2.grad, non_blocking=True)
                return float(l.item())
        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_loss = e2e_step()
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / e2e_steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": B / dt.item(), "unit": UNIT, "ms_per_step": dt.item() * 1e3, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": e2e_steps,
               "entry": "vpa_infonce_step_host (C-ABI, pinned host buffers)" if world == 1 else
                        "vipant_b200.infonce_loss + backward (pinned host tensors, per rank)",
               "loss": e2e_loss}

    if rank == 0:
        pk = peaks(clocks)
        bwd_ms, bwd_n = prof["bwd_sweep"]
        fwd_ms, fwd_n = prof["fwd_sweep"]
        nrm_ms, nrm_n = prof["normalize"]
        roofline = None
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")      # dram bytes per launch from the committed ncu capture
        if os.path.exists(tpath) and B == 32768 and D == 512 and world == 1:
            with open(tpath) as fr:
                traffic = json.load(fr).get("bwd_sweep_dram_bytes")
        if bwd_n:
            flops = 6.0 * b * B * D                         # S recompute + dX1 + dX2 contractions
            ach = flops / (bwd_ms / bwd_n * 1e-3) / 1e12
            roofline = {"kernel": "pair_kernel<BWD> (tcgen05 cta_group::2)", "bound": "tensor", "achieved": ach, "peak": pk["tflops"],
                        "unit": "TFLOP/s", "frac": ach / pk["tflops"], "frac_burst": ach / pk["burst"],
                        "frac_sustained": ach / pk["sustained"], "traffic": traffic, "peak_source": pk["source"],
                        "algorithmic_flops_per_launch": flops, "avg_launch_ms": bwd_ms / bwd_n, "launches": bwd_n}
        step_flops = 8.0 * b * B * D
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": f"InfoNCE fwd+bwd (CELossHead), global batch {B} x {D}, {args.precision} mode, "
                                   f"rows sharded over {world} rank(s)", "global_batch": B, "dim": D, "rows_per_rank": b,
                       "parallelism": f"row-shard x{world}" + (f" + {transport} exchange" if world > 1 else ""),
                       "transport": transport,
                       "l2": "no explicit flush: inputs + operands + partials per step exceed the 126 MB L2",
                       "seed": SEED, "rho": RHO},
            "loss": loss_val, "host_enqueue_ms_per_step": host_ms,
            "step_tflops_algorithmic_8B2D": step_flops / (ms_step * 1e-3) / 1e12,
            "step_frac_of_peak": step_flops / (ms_step * 1e-3) / 1e12 / pk["tflops"],
            "kernel_ms": {"normalize_pair": nrm_ms / max(nrm_n, 1), "fwd_sweep": fwd_ms / max(fwd_n, 1),
                          "bwd_sweep": bwd_ms / max(bwd_n, 1),
                          "fwd_general_gated_off": prof["fwd_general_gated"][0] / max(prof["fwd_general_gated"][1], 1),
                          "finalize_bwd": prof["finalize"][0] / max(prof["finalize"][1], 1),
                          "standalone_relay": prof["push"][0] / max(prof["push"][1], 1)},
            "kernel_ms_note": f"CUDA events around each kernel on the launch stream, on every {PROF_EVERY}th step of the timed region",
            "finalize_hbm": {"achieved_gbs": (2 * b * D * (4 + 4 + 4)) / (prof["finalize"][0] / max(prof["finalize"][1], 1) * 1e-3) / 1e9
                             if prof["finalize"][1] else None, "peak_gbs": pk["hbm"]},
            "normalize_hbm": {"achieved_gbs": (2 * b * D * (4 + 2) + 12 * b) / (nrm_ms / max(nrm_n, 1) * 1e-3) / 1e9 if nrm_n else None,
                              "peak_gbs": pk["hbm"]},
            "roofline": roofline, "clocks": clocks, "e2e": e2e, "parity": parity,
            # kernels of libvipant_b200.so inside the timed region, counted by the library itself (vpa_launch_count).  Per step:
            # normalise, single-pass forward (its first CTAs are the operand all-gather on the peer-memory transport), exact
            # forward (device-gated), [column reduce when b > 8192], statistics (pack + merge; one kernel over peer memory),
            # backward sweep, finalize (+ d logit_scale exchange)
            "gpu_launches": launches, "gpu_launches_per_step": launches / args.steps,
        }
        if world == 1 and not args.no_e2e:
            try:
                line["c2_batch512_latency"] = c2_latency(lib, dev)
            except Exception as exc:
                line["c2_batch512_latency"] = {"error": str(exc)[:200]}
        if world == 1 and not args.no_eager:
            try:
                line["eager_b200"] = eager_b200(dev, D)
            except Exception as exc:
                line["eager_b200"] = {"error": str(exc)[:200]}
        if not args.no_cpu_baseline and world == 1:
            cx1, cx2, cls_, cb, cores = cpu_block_sample(B, D)
            run_cpu_steps(cx1, cx2, cls_, cb, 1)
            times = run_cpu_steps(cx1, cx2, cls_, cb, 4)
            v = cb / float(np.mean(times))
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "host_cpus": os.cpu_count(),
                                    "sample": f"fixed row block of {cb} of the {B} global rows, 4 steps after 1 warm-up (torch "
                                              f"fp32 CPU port of loss_head.py:271-283 + autograd; 12*b*B*D flops = b/B of a "
                                              f"full step), {cores} torch threads"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        sys.stderr.write(f"bench.py: PARITY FAILED on rank {rank}: {parity}\n")
        sys.exit(3)


if __name__ == "__main__":
    main()
