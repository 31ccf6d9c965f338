"""Tensor-level entry points of the B200 InfoNCE / retrieval-scoring path.

PyTorch is used for device memory, streams and ``torch.distributed`` only; all arithmetic runs in
libvipant_b200.so through the C-ABI (``_cabi``).  No CPU / eager fallback exists: CPU tensors raise.

Reference semantics (``/root/reference/cvap/module/decoder/loss_head.py``):
  infonce_loss   <- CELossHead.forward :271-283 (+ autograd), global-batch semantics of the
                    reference's `dp` mode when a process group is given
  l2_normalize   <- :38-40 / :271-273
  sim_rank_topk  <- the `x1s @ x2s.t()` + argsort + where pattern of :109-170, :79-107, :381-385
"""
from __future__ import annotations

import ctypes
import math
import os
from typing import Optional

import torch
import torch.distributed as dist

from . import _cabi

_DTYPES = {torch.float32: _cabi.F32, torch.bfloat16: _cabi.BF16, torch.float16: _cabi.F16}
PRECISIONS = {"bf16": _cabi.PREC_BF16_TC, "fp32": _cabi.PREC_FP32_SIMT}


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _cabi.VipantB200Error(
                "vipant_b200 runs on an sm_100 CUDA device only (got a CPU tensor); there is no CPU fallback")


def _rows2d(x: torch.Tensor) -> torch.Tensor:
    """(rows, D) view the kernels can read in place: unit column stride, 16-byte aligned rows.  Row-strided views
    (``feats[:, :512]``, the CLS token ``hidden[:, 0, :]``) pass through uncopied -- the kernels take a leading dimension."""
    if x.dim() != 2:
        raise ValueError(f"expected a (rows, D) matrix, got shape {tuple(x.shape)}")
    if x.dtype not in _DTYPES:
        x = x.float()
    esz = x.element_size()
    if (x.stride(1) != 1 or x.stride(0) < x.shape[1] or (x.stride(0) * esz) % 16 != 0 or x.data_ptr() % 16 != 0):
        x = x.contiguous()
    return x


def _grad_like(x: torch.Tensor) -> torch.Tensor:
    """Gradient buffer with the SAME leading dimension as x: the backward kernels address x and dx with one `ld`
    (include/vipant_b200.h: "dx1/dx2 same dtype and ld"); `empty_like` would densify a row-strided view."""
    g = torch.empty_strided(x.shape, x.stride(), dtype=x.dtype, device=x.device)
    assert g.stride(0) == x.stride(0) and g.stride(1) == 1
    return g


def _resolve_precision(precision: str, D: int) -> int:
    if precision not in PRECISIONS:
        raise ValueError(f"precision must be one of {sorted(PRECISIONS)}, got {precision!r}")
    return PRECISIONS[precision]


_SCALARS = {}


def _device_scalar(value: torch.Tensor, dev) -> torch.Tensor:
    """A constant scalar (the `scaling=False` temperature, a plain CPU tensor in the reference: loss_head.py:252) on the
    device, cached by value: no blocking pageable H2D copy on every step."""
    key = (float(value), dev.type, dev.index)
    t = _SCALARS.get(key)
    if t is None:
        if len(_SCALARS) > 64:
            _SCALARS.clear()
        t = _SCALARS[key] = torch.tensor(key[0], dtype=torch.float32, device=dev)
    return t


def tensor_core_supported(D: int) -> bool:
    kb = D // 64
    return D % 64 == 0 and 64 <= D <= 512 and (kb <= 4 or kb % 2 == 0)


def l2_normalize(x: torch.Tensor, already_normalized: bool = False) -> torch.Tensor:
    """fp32 ``x / x.norm(dim=-1, keepdim=True)`` (no eps) via the fused normalise kernel."""
    _require_cuda(x)
    x = _rows2d(x)
    rows, D = x.shape
    out = torch.empty((rows, D), dtype=torch.float32, device=x.device)
    if rows:
        with torch.cuda.device(x.device):
            _cabi.check(_cabi.lib().vpa_normalize_cast(_ptr(x), _DTYPES[x.dtype], rows, D, x.stride(0),
                                                       int(already_normalized), None, _ptr(out), None, _stream()),
                        "vpa_normalize_cast")
    return out


class _CudaKernels:
    """The compute steps of one InfoNCE step, each a C-ABI call.  This is the only kernel set the package
    ships; `_sharded_forward/_sharded_backward` take it as a parameter solely so that the row-sharding /
    collective plumbing can be exercised on CPU (gloo) by the tests with a checker standing in for it."""

    def normalize_pair(self, x1, x2, normalized, precision):
        lib = _cabi.lib()
        b, D = x1.shape
        dev = x1.device
        tc = precision == _cabi.PREC_BF16_TC
        fdt = torch.bfloat16 if tc else torch.float32
        a = torch.empty((b, D), dtype=fdt, device=dev)
        t = torch.empty((b, D), dtype=fdt, device=dev)
        inv = torch.empty((2, b), dtype=torch.float32, device=dev)
        dcos = torch.empty((b,), dtype=torch.float32, device=dev)
        _cabi.check(lib.vpa_normalize_pair(
            _ptr(x1), _ptr(x2), _DTYPES[x1.dtype], b, D, x1.stride(0), x2.stride(0), int(normalized),
            _ptr(a) if tc else None, _ptr(t) if tc else None, None if tc else _ptr(a), None if tc else _ptr(t),
            _ptr(inv[0]), _ptr(inv[1]), _ptr(dcos), int(tc), _stream()), "vpa_normalize_pair")
        return a, t, inv, dcos

    def forward_sweep(self, a, t, a_all, t_all, row_offset, logit_scale, scale_max, precision, before_part2=None):
        """Enqueue the forward sweeps.  Returns (col_sum, ws): col_sum (VPA_COLSUM_SPLIT, B) holds, in the single-pass
        regime, the column sums over the LOCAL rows (to be all-reduced over the ranks); zeros otherwise.
        `before_part2()` (optional) runs between the single-pass kernel, which only reads a and t_all, and the rest,
        which also reads a_all: the place to wait for an in-flight all-gather of a_all."""
        lib = _cabi.lib()
        b, D = a.shape
        B = a_all.shape[0]
        dev = a.device
        ws_bytes = lib.vpa_infonce_workspace_bytes(b, B, D, precision)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        col_sum = torch.empty((lib.vpa_infonce_colsum_floats(B) // B, B), dtype=torch.float32, device=dev)
        cap = float(scale_max) if scale_max else 0.0                     # 0 -> no clamp (`or inf`, :254)
        for parts in ((1, 2) if before_part2 is not None else (3,)):
            if parts == 2:
                before_part2()
            _cabi.check(lib.vpa_infonce_fwd_sweep(
                _ptr(a), _ptr(t), _ptr(a_all), _ptr(t_all), precision, b, B, D, row_offset, _ptr(logit_scale), cap,
                _ptr(ws), ws_bytes, _ptr(col_sum), parts, _stream()), "vpa_infonce_fwd_sweep")
        return col_sum, ws

    def forward_finish(self, b, B, D, row_offset, logit_scale, scale_max, dcos, precision, ws, col_sum):
        lib = _cabi.lib()
        dev = dcos.device
        stats = torch.empty((3, b), dtype=torch.float32, device=dev)     # row_lse, col_lse, diag
        scale = torch.empty((2,), dtype=torch.float32, device=dev)       # s, grad-flows flag
        cap = float(scale_max) if scale_max else 0.0
        _cabi.check(lib.vpa_infonce_fwd_finish(
            precision, b, B, D, row_offset, _ptr(logit_scale), cap, _ptr(dcos), _ptr(ws), ws.numel(), _ptr(col_sum),
            _ptr(stats[0]), _ptr(stats[1]), _ptr(stats[2]), _ptr(scale), _stream()), "vpa_infonce_fwd_finish")
        return stats, scale

    def loss(self, stats_all):
        B = stats_all.shape[1]
        loss = torch.empty((), dtype=torch.float32, device=stats_all.device)
        _cabi.check(_cabi.lib().vpa_infonce_loss(_ptr(stats_all[0]), _ptr(stats_all[1]), _ptr(stats_all[2]), B,
                                                 _ptr(loss), _stream()), "vpa_infonce_loss")
        return loss

    def backward(self, x1, x2, a, t, a_all, t_all, inv, stats_all, scale, ws, row_offset, grad_out, normalized,
                 precision):
        lib = _cabi.lib()
        b, D = a.shape
        B = a_all.shape[0]
        dev = a.device
        g = grad_out.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        dx1 = _grad_like(x1)
        dx2 = _grad_like(x2)
        dls = torch.empty((), dtype=torch.float32, device=dev)
        _cabi.check(lib.vpa_infonce_bwd(
            _ptr(a), _ptr(t), _ptr(a_all), _ptr(t_all), precision, b, B, D, row_offset, _ptr(scale),
            _ptr(stats_all[0]), _ptr(stats_all[1]), _ptr(g), _ptr(x1), _ptr(x2), _DTYPES[x1.dtype],
            x1.stride(0), x2.stride(0), _ptr(inv[0]), _ptr(inv[1]), int(normalized), _ptr(ws), ws.numel(),
            _ptr(dx1), _ptr(dx2), _ptr(dls), _stream()), "vpa_infonce_bwd")
        return dx1, dx2, dls


_KERNELS = _CudaKernels()


def _sharded_forward(kern, x1, x2, logit_scale, scale_max, normalized, precision, group):
    """Row-sharded global-batch forward: rank r owns rows [r*b, (r+1)*b).

    normalise locally -> all-gather the normalised features -> forward sweeps of the local rows against ALL rows
    -> all-reduce of the (8, B) column sums (the one exchange step; only meaningful in the single-pass regime, where
    each rank holds the sums over its own rows) -> statistics of the local rows -> all-gather of the three (b,)
    statistic vectors -> identical global loss on every rank.
    """
    b, D = x1.shape
    world = dist.get_world_size(group) if group is not None else 1
    rank = dist.get_rank(group) if group is not None else 0
    a, t, inv, dcos = kern.normalize_pair(x1, x2, normalized, precision)
    wait_a = None
    if world > 1:
        a_all = torch.empty((b * world, a.shape[1]), dtype=a.dtype, device=a.device)
        t_all = torch.empty_like(a_all)
        # x2 operands first: the single-pass forward needs only them; the gather of the x1 operands overlaps with it
        work_t = dist.all_gather_into_tensor(t_all, t, group=group, async_op=True)
        work_a = dist.all_gather_into_tensor(a_all, a, group=group, async_op=True)
        work_t.wait()
        wait_a = work_a.wait
    else:
        a_all, t_all = a, t
    col_sum, ws = kern.forward_sweep(a, t, a_all, t_all, rank * b, logit_scale, scale_max, precision, before_part2=wait_a)
    if world > 1:
        dist.all_reduce(col_sum, group=group)
    stats, scale = kern.forward_finish(b, b * world, D, rank * b, logit_scale, scale_max, dcos, precision, ws, col_sum)
    if world > 1:
        gathered = torch.empty((world * 3, b), dtype=stats.dtype, device=stats.device)
        dist.all_gather_into_tensor(gathered, stats.contiguous(), group=group)
        stats_all = gathered.view(world, 3, b).permute(1, 0, 2).reshape(3, b * world).contiguous()
    else:
        stats_all = stats
    loss = kern.loss(stats_all)
    saved = (x1, x2, a, t, a_all, t_all, inv, stats_all, scale)
    return loss, saved, (ws, rank * b, world)


def _sharded_backward(kern, saved, extra, grad_out, normalized, precision, group, dls_sum=True):
    """dL/d(local rows): two local sweeps (x1 rows vs all x2, x2 rows vs all x1); only the replicated
    logit_scale gradient needs a (scalar) all-reduce (skipped when the caller wants this rank's partial)."""
    x1, x2, a, t, a_all, t_all, inv, stats_all, scale = saved
    ws, row_offset, world = extra
    dx1, dx2, dls = kern.backward(x1, x2, a, t, a_all, t_all, inv, stats_all, scale, ws, row_offset, grad_out,
                                  normalized, precision)
    if world > 1 and dls_sum:
        dist.all_reduce(dls, group=group)
    return dx1, dx2, dls


class _InfoNCEFunction(torch.autograd.Function):
    """Host-orchestrated step (a dozen C-ABI calls + torch.distributed collectives).  Kept for process groups that are
    not NCCL and as the reference orchestration the gloo tests exercise; CUDA + NCCL / single GPU use _FusedStep."""

    @staticmethod
    def forward(ctx, x1, x2, logit_scale, scale_max, normalized, precision, group, dls_sum, seg_key):
        with torch.cuda.device(x1.device):
            loss, saved, extra = _sharded_forward(_KERNELS, x1, x2, logit_scale, scale_max, normalized, precision, group)
        ctx.save_for_backward(*saved)
        ctx.cfg = (extra, normalized, precision, group, dls_sum)
        ctx.set_materialize_grads(False)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        if grad_out is None:
            return (None,) * 9
        extra, normalized, precision, group, dls_sum = ctx.cfg
        with torch.cuda.device(ctx.saved_tensors[0].device):
            dx1, dx2, dls = _sharded_backward(_KERNELS, ctx.saved_tensors, extra, grad_out, normalized, precision, group,
                                              dls_sum)
        return (dx1, dx2, dls) + (None,) * 6


# ---------------------------------------------------------------------------- in-library orchestration (2 calls / step)
_COMMS = {}


def _nccl_path():
    try:
        import nvidia.nccl as _n
        for base in list(_n.__path__):
            cand = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                return cand
    except Exception:
        pass
    return None


def _get_comm(group, dev):
    """NCCL communicator of the library for (`group`, device) (created once; call with `dev` current): rank 0's unique id is
    broadcast with torch.distributed.  The cache holds the group object, so its id cannot be recycled."""
    key = (id(group), dev.index)
    if key in _COMMS:
        return _COMMS[key][:3]
    lib = _cabi.lib()
    path = _nccl_path()
    _cabi.check(lib.vpa_comm_load(path.encode() if path else None), "vpa_comm_load")
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = ctypes.create_string_buffer(128)
        _cabi.check(lib.vpa_comm_unique_id(buf), "vpa_comm_unique_id")
        uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    uid = uid.to(dev)
    dist.broadcast(uid, src=dist.get_global_rank(group, 0) if hasattr(dist, "get_global_rank") else 0, group=group)
    raw = bytes(uid.cpu().numpy().tobytes())
    comm = ctypes.c_void_p()
    _cabi.check(lib.vpa_comm_init(raw, rank, world, ctypes.byref(comm)), "vpa_comm_init")
    _COMMS[key] = (comm, rank, world, group)
    return _COMMS[key][:3]


class _FusedStep(torch.autograd.Function):
    """vpa_infonce_fwd_sharded / vpa_infonce_bwd_sharded: the whole step, collectives included, in two C-ABI calls."""

    @staticmethod
    def forward(ctx, x1, x2, logit_scale, scale_max, normalized, precision, group, dls_sum, seg_key):
        lib = _cabi.lib()
        b, D = x1.shape
        dev = x1.device
        with torch.cuda.device(dev):
            comm, rank, world = (None, 0, 1) if group is None else _get_comm(group, dev)
            nbytes = lib.vpa_sharded_state_bytes(b, world, D, precision)
            state = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            cap = float(scale_max) if scale_max else 0.0
            _cabi.check(lib.vpa_infonce_fwd_sharded(
                comm, _ptr(x1), _ptr(x2), _DTYPES[x1.dtype], b, world, rank, D, x1.stride(0), x2.stride(0), int(normalized),
                _ptr(logit_scale), cap, precision, _ptr(state), nbytes, _ptr(loss), _stream()), "vpa_infonce_fwd_sharded")
        ctx.save_for_backward(x1, x2, state)
        ctx.cfg = (comm, rank, world, normalized, precision, dls_sum)
        ctx.set_materialize_grads(False)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        if grad_out is None:
            return (None,) * 9
        comm, rank, world, normalized, precision, dls_sum = ctx.cfg
        x1, x2, state = ctx.saved_tensors
        lib = _cabi.lib()
        dev = x1.device
        b, D = x1.shape
        with torch.cuda.device(dev):
            g = grad_out.detach().to(device=dev, dtype=torch.float32).reshape(1)
            dx1 = _grad_like(x1)
            dx2 = _grad_like(x2)
            dls = torch.empty((), dtype=torch.float32, device=dev)
            _cabi.check(lib.vpa_infonce_bwd_sharded(
                comm, _ptr(x1), _ptr(x2), _DTYPES[x1.dtype], b, world, rank, D, x1.stride(0), x2.stride(0), int(normalized),
                precision, _ptr(g), _ptr(state), state.numel(), _ptr(dx1), _ptr(dx2), _ptr(dls), int(dls_sum), _stream()),
                "vpa_infonce_bwd_sharded")
        return (dx1, dx2, dls) + (None,) * 6


# ---------------------------------------------------------------------------- peer-memory transport (no NCCL on the data path)
_P2P = {}


def _get_p2p(group, b, D, precision, dev, seg_key=None):
    """Symmetric segment of the library for (group, shape, call site): created once; the 64-byte CUDA IPC handles are
    exchanged with torch.distributed (any backend: object all-gather), then every rank maps its peers' segments.  A segment
    keeps the forward state of its two most recent steps, so every call site that runs within one training step (the pairs
    of a composite head) owns one: `seg_key` tells them apart.  Returns None when some rank cannot set it up (no peer
    access / IPC): the ranks agree on that and the caller uses the NCCL transport."""
    key = (id(group), b, D, precision, dev.index, seg_key)
    if key in _P2P:
        return _P2P[key]
    lib = _cabi.lib()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    handle = ctypes.c_void_p()
    mine = ctypes.create_string_buffer(64)
    rc = lib.vpa_p2p_create(b, world, rank, D, precision, ctypes.byref(handle), mine)
    err = "" if rc == 0 else (lib.vpa_last_error_string() or b"").decode()
    everyone = [None] * world
    dist.all_gather_object(everyone, (rc, bytes(mine.raw), os.getpid()), group=group)
    ok = all(r == 0 for r, _, _ in everyone)
    if ok:
        rc = lib.vpa_p2p_connect(handle, b"".join(h for _, h, _ in everyone))
        err = "" if rc == 0 else (lib.vpa_last_error_string() or b"").decode()
    status = [None] * world
    dist.all_gather_object(status, rc if ok else 1, group=group)    # also: nobody stores into a peer before all have mapped
    if not all(r == 0 for r in status):
        if handle:
            lib.vpa_p2p_destroy(handle)
        if os.environ.get("VIPANT_REQUIRE_P2P"):          # tests / A-B runs: no silent change of transport
            raise _cabi.VipantB200Error(f"peer-memory transport unavailable on rank {rank}: {err or 'a peer failed'}")
        import warnings
        warnings.warn(f"vipant_b200: peer-memory transport unavailable on rank {rank} ({err or 'a peer failed'}); using NCCL")
        _P2P[key] = None
        return None
    _P2P[key] = (handle, rank, world, group)      # (holds the group: its id cannot be recycled while the entry lives)
    if not _P2P_ATEXIT:
        import atexit
        atexit.register(_destroy_p2p)
        _P2P_ATEXIT.append(True)
    return _P2P[key]


_P2P_ATEXIT = []


def _destroy_p2p():
    lib = _cabi.lib()
    for entry in _P2P.values():
        if entry is not None:
            lib.vpa_p2p_destroy(entry[0])
    _P2P.clear()


class _P2PStep(torch.autograd.Function):
    """vpa_infonce_fwd_p2p / vpa_infonce_bwd_p2p: the row-sharded step with every exchange done by kernels storing into
    the peers' memory (all-gather fused with the forward sweep); two C-ABI calls per step."""

    @staticmethod
    def forward(ctx, x1, x2, logit_scale, scale_max, normalized, precision, group, dls_sum, seg_key):
        lib = _cabi.lib()
        b, D = x1.shape
        dev = x1.device
        with torch.cuda.device(dev):
            handle, rank, world = _get_p2p(group, b, D, precision, dev, seg_key)[:3]
            loss = torch.empty((), dtype=torch.float32, device=dev)
            epoch = ctypes.c_uint32()
            cap = float(scale_max) if scale_max else 0.0
            _cabi.check(lib.vpa_infonce_fwd_p2p(
                handle, _ptr(x1), _ptr(x2), _DTYPES[x1.dtype], b, world, rank, D, x1.stride(0), x2.stride(0), int(normalized),
                _ptr(logit_scale), cap, precision, _ptr(loss), ctypes.byref(epoch), _stream()), "vpa_infonce_fwd_p2p")
        ctx.save_for_backward(x1, x2)
        ctx.cfg = (handle, epoch.value, rank, world, normalized, precision, dls_sum)
        ctx.set_materialize_grads(False)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        if grad_out is None:
            return (None,) * 9
        handle, epoch, rank, world, normalized, precision, dls_sum = ctx.cfg
        x1, x2 = ctx.saved_tensors
        lib = _cabi.lib()
        dev = x1.device
        b, D = x1.shape
        with torch.cuda.device(dev):
            g = grad_out.detach().to(device=dev, dtype=torch.float32).reshape(1)
            dx1 = _grad_like(x1)
            dx2 = _grad_like(x2)
            dls = torch.empty((), dtype=torch.float32, device=dev)
            _cabi.check(lib.vpa_infonce_bwd_p2p(
                handle, epoch, _ptr(x1), _ptr(x2), _DTYPES[x1.dtype], b, world, rank, D, x1.stride(0), x2.stride(0),
                int(normalized), precision, _ptr(g), _ptr(dx1), _ptr(dx2), _ptr(dls), int(dls_sum), _stream()),
                "vpa_infonce_bwd_p2p")
        return (dx1, dx2, dls) + (None,) * 6


def _transport(group) -> str:
    """"local" (one GPU), "p2p" (peer-memory kernels, default for <= 8 ranks of one node), "nccl" (in-library NCCL
    orchestration) or "host" (torch.distributed collectives between C-ABI calls).  VIPANT_TRANSPORT overrides."""
    if os.environ.get("VIPANT_HOST_ORCHESTRATION"):          # A/B knob: the host-orchestrated path
        return "host"
    if group is None or dist.get_world_size(group) == 1:
        return "local"
    forced = os.environ.get("VIPANT_TRANSPORT")
    if forced in ("p2p", "nccl", "host"):
        return forced
    if dist.get_world_size(group) <= 8:
        return "p2p"
    try:
        return "nccl" if dist.get_backend(group) == "nccl" else "host"
    except Exception:
        return "host"


def infonce_loss(x1: torch.Tensor, x2: torch.Tensor, logit_scale: torch.Tensor, scale_max=None,
                 normalized: bool = False, precision: str = "bf16",
                 group: Optional["dist.ProcessGroup"] = None, logit_scale_grad: str = "sum",
                 segment_key=None) -> torch.Tensor:
    """Symmetric InfoNCE  CE(s*a@t.T, arange) + CE(s*t@a.T, arange)  with s = min(exp(logit_scale), scale_max).

    x1, x2: (b, D) CUDA tensors (fp32 / bf16 / fp16), this process's rows.  With ``group`` the loss is the
    GLOBAL-batch loss over the concatenation of all ranks' rows in rank order (each rank must pass the same b);
    the returned feature gradients are d(global loss)/d(local rows).
    logit_scale_grad: "sum" -- d/d logit_scale is summed over the ranks (the full derivative, identical everywhere; for
    trainers that do NOT reduce the parameter's gradient again); "local" -- this rank's partial (sum over its own rows),
    like the feature gradients: what a DistributedDataParallel trainer wants (DDP averages both kinds, so multiplying the
    loss by the world size then reproduces the single-process update of the reference's `dp` mode exactly).
    segment_key: any hashable that identifies the call site; call sites that run within one training step (the pairs of a
    composite head) must use different keys -- each owns a peer-memory segment holding its forward state.
    precision "bf16": tcgen05 tensor cores (needs D in {64,128,192,256,384,512}); "fp32": exact FFMA path.
    """
    if logit_scale_grad not in ("sum", "local"):
        raise ValueError(f"logit_scale_grad must be 'sum' or 'local', got {logit_scale_grad!r}")
    _require_cuda(x1, x2)
    x1, x2 = _rows2d(x1), _rows2d(x2)
    if x1.shape != x2.shape:
        raise ValueError(f"x1 {tuple(x1.shape)} and x2 {tuple(x2.shape)} must have the same shape")
    if x1.dtype != x2.dtype:
        x1, x2 = x1.float(), x2.float()
    if x1.shape[0] == 0:
        raise ValueError("empty batch")
    prec = _resolve_precision(precision, x1.shape[1])
    if prec == _cabi.PREC_BF16_TC and not tensor_core_supported(x1.shape[1]):
        raise _cabi.VipantB200Error(f"precision='bf16' needs D in {{64,128,192,256,384,512}}, got D={x1.shape[1]}; "
                                    "use precision='fp32'")
    ls = logit_scale
    if not isinstance(ls, torch.Tensor):
        ls = torch.tensor(float(ls))
    if ls.device != x1.device or ls.dtype != torch.float32:
        # `scaling=False` heads keep a plain CPU tensor (loss_head.py:252); its value is copied, no grad needed
        if ls.device.type == "cpu" and not ls.requires_grad:
            ls = _device_scalar(ls, x1.device)
        else:
            ls = ls.to(device=x1.device, dtype=torch.float32)
    transport = _transport(group)
    if transport == "local":
        group = None
    if transport == "p2p":
        with torch.cuda.device(x1.device):
            if _get_p2p(group, x1.shape[0], x1.shape[1], prec, x1.device, segment_key) is None:
                transport = "nccl" if dist.get_backend(group) == "nccl" else "host"
    fn = {"local": _FusedStep, "nccl": _FusedStep, "p2p": _P2PStep, "host": _InfoNCEFunction}[transport]
    return fn.apply(x1, x2, ls.reshape(()), scale_max, bool(normalized), prec, group, logit_scale_grad == "sum", segment_key)


# ---------------------------------------------------------------------------- several pairs over shared modalities
class _MultiPairStep(torch.autograd.Function):
    """vpa_infonce_multi_fwd / _bwd: up to five InfoNCE pairs over up to five feature matrices in one set of launches."""

    @staticmethod
    def forward(ctx, pairs, scale_maxes, normalized, precision, n_mod, *tensors):
        lib = _cabi.lib()
        feats, scales = tensors[:n_mod], tensors[n_mod:]
        rows, D = feats[0].shape
        dev = feats[0].device
        n_pairs = len(pairs)
        with torch.cuda.device(dev):
            nbytes = lib.vpa_infonce_multi_state_bytes(rows, D, n_mod, n_pairs, precision)
            state = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
            losses = torch.empty((n_pairs,), dtype=torch.float32, device=dev)
            xp = (ctypes.c_void_p * n_mod)(*[f.data_ptr() for f in feats])
            ld = (ctypes.c_int64 * n_mod)(*[f.stride(0) for f in feats])
            px = (ctypes.c_int32 * n_pairs)(*[p[0] for p in pairs])
            py = (ctypes.c_int32 * n_pairs)(*[p[1] for p in pairs])
            lsp = (ctypes.c_void_p * n_pairs)(*[t.data_ptr() for t in scales])
            caps = (ctypes.c_float * n_pairs)(*[float(c) if c else 0.0 for c in scale_maxes])
            _cabi.check(lib.vpa_infonce_multi_fwd(xp, ld, _DTYPES[feats[0].dtype], rows, D, n_mod, int(normalized), px, py, n_pairs,
                                                  lsp, caps, precision, _ptr(state), nbytes, _ptr(losses), _stream()),
                        "vpa_infonce_multi_fwd")
        ctx.save_for_backward(state, *feats)
        ctx.cfg = (pairs, normalized, precision, n_mod)
        return losses

    @staticmethod
    def backward(ctx, grad_losses):
        pairs, normalized, precision, n_mod = ctx.cfg
        state, *feats = ctx.saved_tensors
        lib = _cabi.lib()
        rows, D = feats[0].shape
        dev = feats[0].device
        n_pairs = len(pairs)
        with torch.cuda.device(dev):
            g = grad_losses.detach().to(device=dev, dtype=torch.float32).contiguous()
            dxs = [_grad_like(f) for f in feats]
            dls = torch.empty((n_pairs,), dtype=torch.float32, device=dev)
            xp = (ctypes.c_void_p * n_mod)(*[f.data_ptr() for f in feats])
            ld = (ctypes.c_int64 * n_mod)(*[f.stride(0) for f in feats])
            dp = (ctypes.c_void_p * n_mod)(*[d.data_ptr() for d in dxs])
            px = (ctypes.c_int32 * n_pairs)(*[p[0] for p in pairs])
            py = (ctypes.c_int32 * n_pairs)(*[p[1] for p in pairs])
            _cabi.check(lib.vpa_infonce_multi_bwd(xp, ld, _DTYPES[feats[0].dtype], rows, D, n_mod, int(normalized), px, py, n_pairs,
                                                  precision, _ptr(g), _ptr(state), state.numel(), dp, _ptr(dls), _stream()),
                        "vpa_infonce_multi_bwd")
        return (None, None, None, None, None) + tuple(dxs) + tuple(dls[k] for k in range(n_pairs))


def multi_pair_supported(feats, precision: str = "bf16") -> bool:
    """Shapes the fused multi-pair step covers: CUDA matrices of one shape (rows <= 8192, D in {256, 512}) and dtype."""
    if precision != "bf16" or not feats:
        return False
    f0 = feats[0]
    return all(isinstance(f, torch.Tensor) and f.is_cuda and f.dim() == 2 and f.shape == f0.shape and f.dtype == f0.dtype
               for f in feats) and f0.shape[1] in (256, 512) and 0 < f0.shape[0] <= 8192


def infonce_multi_loss(feats, pairs, logit_scales, scale_maxes=None, normalized: bool = False, precision: str = "bf16"):
    """Per-pair symmetric InfoNCE losses of several pairs over shared feature matrices (the composite heads of the reference,
    loss_head.py:421-598) in ONE fused step: every matrix is normalised once, all pairs share the sweep launches, and the
    gradient of a matrix is the sum over the pairs it takes part in.

    feats: list of (rows, D) CUDA tensors (same shape / dtype); pairs: list of (i, j) indices into feats; logit_scales: one
    0-d fp32 CUDA tensor per pair.  Returns a (len(pairs),) tensor of losses, differentiable w.r.t. feats and logit_scales.
    """
    feats = [_rows2d(f) for f in feats]
    if not multi_pair_supported(feats, precision):
        raise _cabi.VipantB200Error("infonce_multi_loss: needs CUDA matrices of one shape with rows <= 8192 and D in {256, 512} "
                                    "(precision 'bf16'); run the pairs one by one with infonce_loss otherwise")
    if not 1 <= len(pairs) <= 5 or len(feats) > 5 or len(logit_scales) != len(pairs):
        raise ValueError("infonce_multi_loss: 1..5 pairs over at most 5 matrices, one logit_scale per pair")
    dev = feats[0].device
    scales = []
    for ls in logit_scales:
        if not isinstance(ls, torch.Tensor):
            ls = torch.tensor(float(ls))
        if ls.device != dev or ls.dtype != torch.float32:
            ls = _device_scalar(ls, dev) if (ls.device.type == "cpu" and not ls.requires_grad) else ls.to(device=dev, dtype=torch.float32)
        scales.append(ls.reshape(()))
    caps = list(scale_maxes) if scale_maxes is not None else [None] * len(pairs)
    return _MultiPairStep.apply(tuple((int(i), int(j)) for i, j in pairs), tuple(caps), bool(normalized), PRECISIONS[precision],
                                len(feats), *feats, *scales)


def _gt_matrix(gt, rows, upper, dev, what):
    """(rows,) or (rows, g <= 8) integer ground-truth indices -> int32 (rows, g), validated on the host side of the call:
    an index outside [0, upper) would make the kernel compare against another row's memory."""
    if gt is None:
        return None, 0
    gt2 = gt[:, None] if gt.dim() == 1 else gt           # rows may be 0
    if gt2.dim() != 2 or gt2.shape[0] != rows:
        raise ValueError(f"{what} must have shape (N,) or (N, g) with N={rows}, got {tuple(gt.shape)}")
    if gt2.shape[1] > 8:
        raise ValueError(f"{what}: at most 8 ground-truth indices per row, got {gt2.shape[1]}")
    g32 = gt2.to(device=dev, dtype=torch.int32).contiguous()
    return g32, g32.shape[1]


def check_gt_range(gt: torch.Tensor, upper: int, what: str = "gt") -> None:
    """Raise if an index lies outside [0, upper).  One device->host sync: call it once per evaluation, not per batch."""
    if gt is not None and gt.numel() and (int(gt.min()) < 0 or int(gt.max()) >= upper):
        raise ValueError(f"{what}: indices must lie in [0, {upper})")


def sim_rank_fused(q: torch.Tensor, k: torch.Tensor, gt_q: Optional[torch.Tensor] = None,
                   gt_k: Optional[torch.Tensor] = None, top1_q: bool = False, top1_k: bool = False):
    """Both directions of the monitors' scoring from ONE pass over the similarity ``q @ k.T`` (never written to memory).

    q (N, D), k (M, D): CUDA, converted to fp32.  gt_q: (N,) / (N, g<=8) key indices whose rank within each query's row
    is wanted; gt_k: (M,) / (M, g<=8) query indices whose rank within each key's row of ``k @ q.T`` is wanted.
    Returns a dict: ranks_q (N, g) int64, ranks_k (M, g) int64, top1_q / top1_k (idx int64, val fp32) -- entries only for
    what was asked.  rank = 0-based position in a stable descending sort; NaN similarities sort first (as torch.argsort
    does); out-of-range ground-truth indices give rank 0 -- validate with check_gt_range when they come from outside.
    """
    _require_cuda(q, k, gt_q, gt_k)
    q = _rows2d(q.float())
    k = _rows2d(k.float())
    N, D = q.shape
    M = k.shape[0]
    if k.shape[1] != D:
        raise ValueError(f"feature dims differ: {D} vs {k.shape[1]}")
    dev = q.device
    gq, ng_q = _gt_matrix(gt_q, N, M, dev, "gt_q")
    gk, ng_k = _gt_matrix(gt_k, M, N, dev, "gt_k")
    out = {}
    rq = torch.zeros((N, ng_q), dtype=torch.int32, device=dev) if ng_q else None
    rk = torch.zeros((M, ng_k), dtype=torch.int32, device=dev) if ng_k else None
    iq = torch.full((N,), -1, dtype=torch.int64, device=dev) if top1_q else None
    vq = torch.empty((N,), dtype=torch.float32, device=dev) if top1_q else None
    ik = torch.full((M,), -1, dtype=torch.int64, device=dev) if top1_k else None
    vk = torch.empty((M,), dtype=torch.float32, device=dev) if top1_k else None
    if N and M:
        lib = _cabi.lib()
        with torch.cuda.device(dev):
            ws_bytes = lib.vpa_sim_fused_workspace_bytes(N, M, ng_q, ng_k)
            ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
            _cabi.check(lib.vpa_sim_rank_fused(_ptr(q), _ptr(k), N, M, D, q.stride(0), k.stride(0), _ptr(gq), ng_q, _ptr(gk),
                                               ng_k, _ptr(rq), _ptr(rk), _ptr(iq), _ptr(vq), _ptr(ik), _ptr(vk), _ptr(ws),
                                               ws_bytes, _stream()), "vpa_sim_rank_fused")
    if rq is not None:
        out["ranks_q"] = rq.long()
    if rk is not None:
        out["ranks_k"] = rk.long()
    if top1_q:
        out["top1_q"] = (iq, vq)
    if top1_k:
        out["top1_k"] = (ik, vk)
    return out


def sim_rank_topk(q: torch.Tensor, k: torch.Tensor, gt: Optional[torch.Tensor] = None, topk: int = 0):
    """Similarity ``q @ k.T`` (fp32) reduced on the fly to ranks of ground-truth columns and the top-k keys.

    q (N, D), k (M, D): CUDA, converted to fp32.  gt: (N,) or (N, g<=8) integer column indices.
    Returns (ranks int64 (N, g) or None, topk_idx int64 (N, topk) or None, topk_val fp32 or None);
    rank = 0-based position of the column in a stable descending sort of the row.
    topk <= 1 runs the fused kernel (the similarity is consumed in registers); topk > 1 materialises S in a workspace.
    NaN similarities: the fused path sorts them first (like torch.argsort(descending=True)), the topk > 1 path last.
    """
    if topk <= 1:
        if q.dim() == 2 and k.dim() == 2 and topk > k.shape[0]:
            raise ValueError("topk exceeds the number of keys")
        res = sim_rank_fused(q, k, gt_q=gt, top1_q=topk == 1)
        idx = val = None
        if topk == 1:
            idx, val = res["top1_q"][0][:, None], res["top1_q"][1][:, None]
        return res.get("ranks_q"), idx, val
    _require_cuda(q, k, gt)
    q = _rows2d(q.float())
    k = _rows2d(k.float())
    N, D = q.shape
    M = k.shape[0]
    if k.shape[1] != D:
        raise ValueError(f"feature dims differ: {D} vs {k.shape[1]}")
    dev = q.device
    gt32, g = _gt_matrix(gt, N, M, dev, "gt")
    if gt32 is not None:
        check_gt_range(gt32, M)          # this kernel reads S[gt] unchecked
    ranks = torch.empty((N, g), dtype=torch.int32, device=dev) if g else None
    idx = torch.empty((N, topk), dtype=torch.int64, device=dev) if topk else None
    val = torch.empty((N, topk), dtype=torch.float32, device=dev) if topk else None
    if N:
        lib = _cabi.lib()
        with torch.cuda.device(dev):
            ws_bytes = lib.vpa_sim_workspace_bytes(N, M)
            ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
            _cabi.check(lib.vpa_sim_rank_topk(_ptr(q), _ptr(k), N, M, D, q.stride(0), k.stride(0), _ptr(gt32), g, topk,
                                              _ptr(idx), _ptr(val), _ptr(ranks), _ptr(ws), ws_bytes, _stream()),
                        "vpa_sim_rank_topk")
    return (ranks.long() if ranks is not None else None), idx, val
