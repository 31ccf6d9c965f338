"""Row-sharded global-batch InfoNCE over NCCL (one process per GPU) against the single-process oracle.

The reference computes the loss on the gathered global batch (`dp` mode); under one process per GPU
the same value is obtained by: local normalise -> all-gather of the bf16 operands -> row/column statistics of the
local rows against ALL rows -> all-gather of the (b,) statistics -> local backward sweeps -> scalar all-reduce of
d logit_scale.  Needs >= 2 GPUs (skipped otherwise; the plumbing itself is covered on gloo in test_host_logic.py).
"""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import infonce_oracle as io

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, D, precision, transport, shared_device, steps, out, env=None):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), VIPANT_TRANSPORT=transport)
    os.environ.update(env or {})
    dev = 0 if shared_device else rank
    torch.cuda.set_device(dev)
    if shared_device:      # ranks share GPU 0 (NCCL refuses that): gloo only carries the IPC-handle exchange
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
    try:
        import vipant_b200 as vb
        lsv, rho = _case(env)
        x1n, x2n = io.make_pair(B, D, rho, 21)
        b = B // world
        x1 = torch.from_numpy(x1n[rank * b:(rank + 1) * b]).cuda().requires_grad_(True)
        x2 = torch.from_numpy(x2n[rank * b:(rank + 1) * b]).cuda().requires_grad_(True)
        ls = torch.tensor(lsv, device="cuda", requires_grad=True)
        gen = torch.Generator(device="cuda").manual_seed(100 + rank)
        for i in range(steps):       # several steps: epochs, buffer parity and flag reuse of the peer-memory transport
            x1.grad = x2.grad = ls.grad = None
            if i < steps - 1:        # different data on the earlier steps: a stale operand / message buffer would show
                y1 = (x1.detach() + torch.randn(x1.shape, device="cuda", generator=gen)).requires_grad_(True)
                y2 = (x2.detach() + torch.randn(x2.shape, device="cuda", generator=gen)).requires_grad_(True)
                vb.infonce_loss(y1, y2, ls, precision=precision, group=dist.group.WORLD).backward()
                continue
            loss = vb.infonce_loss(x1, x2, ls, precision=precision, group=dist.group.WORLD)
            (loss * 3.0).backward()
        torch.cuda.synchronize()
        out[rank] = (loss.item(), x1.grad.cpu().numpy(), x2.grad.cpu().numpy(), ls.grad.item())
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _case(env):
    """(logit_scale, rho): TEST_SCALE100 selects the exact two-sweep regime (s = 100 > 43) on data whose loss is not tiny."""
    if env and env.get("TEST_SCALE100"):
        return math.log(100.0), 0.1
    return math.log(1 / 0.07), 0.3


def _check(out, world, B, D, precision, env=None):
    lsv, rho = _case(env)
    x1n, x2n = io.make_pair(B, D, rho, 21)
    ref = io.infonce_closed_form(x1n, x2n, float(np.float32(lsv)), grad_output=3.0)
    tl, tg = (1e-3, 1e-2) if precision == "bf16" else (1e-4, 1e-3)
    b = B // world

    def rel(a, r):
        return float(np.linalg.norm(a.astype(np.float64) - r) / np.linalg.norm(r))
    for r in range(world):
        loss, dx1, dx2, dls = out[r]
        assert abs(loss - ref.loss) <= tl * abs(ref.loss)
        assert rel(dx1, ref.dx1[r * b:(r + 1) * b]) <= tg and rel(dx2, ref.dx2[r * b:(r + 1) * b]) <= tg
        assert abs(dls - ref.dlogit_scale) <= tg * abs(ref.dlogit_scale)
    for r in range(1, world):
        assert out[0][0] == out[r][0] and out[0][3] == out[r][3]      # identical global loss / d logit_scale on every rank


# env: VPA_FWD1_CHUNKS=1 makes one CTA pair sweep ALL rank blocks (the chunk-major tile order over several peer blocks);
# VPA_P2P_RELAY_CTAS=2 leaves one relay CTA pair to move every chunk (long item lists, ring wrap-around); D = 128 / 384 and
# fp32 take the stand-alone relay kernel (shapes without the fused forward); TEST_SCALE100 the exact two-sweep regime, in
# which the fused kernel's sweep CTAs return at once and only its relays work
@pytest.mark.parametrize("world,precision,B,D,env", [
    (2, "bf16", 1024, 512, None), (2, "bf16", 600, 256, None), (2, "fp32", 256, 128, None),
    (4, "bf16", 2048, 512, {"VPA_FWD1_CHUNKS": "1"}), (3, "bf16", 1152, 512, None),
    (4, "bf16", 4096, 256, {"VPA_FWD1_CHUNKS": "2", "VPA_P2P_RELAY_CTAS": "2"}),
    (2, "bf16", 1024, 512, {"TEST_SCALE100": "1"}),
    (2, "bf16", 640, 384, None), (3, "bf16", 768, 128, {"TEST_SCALE100": "1"})])
def test_p2p_transport_ranks_sharing_one_gpu(world, precision, B, D, env):
    """The peer-memory transport (CUDA IPC segments, operand transfer + arrival flags consumed by the forward sweep,
    message and d logit_scale exchange by peer stores) between processes that share GPU 0: runs on a single-GPU box."""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), B, D, precision, "p2p", True, 3, out, env), nprocs=world, join=True)
    _check(out, world, B, D, precision, env)


@pytest.mark.parametrize("transport", ["p2p", "nccl", "host"])
@pytest.mark.parametrize("precision,B,D", [("bf16", 1024, 512), ("bf16", 600, 256), ("fp32", 256, 128)])
def test_sharded_matches_global_batch(precision, B, D, transport):
    world = 2
    if torch.cuda.device_count() < world:
        pytest.skip("needs 2 GPUs")
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), B, D, precision, transport, False, 4, out), nprocs=world, join=True)
    _check(out, world, B, D, precision)


def _worker_many_steps(rank, world, port, B, D, steps, out):
    """>= 50 steps with DIFFERENT data every step: shakes the parity buffers / epoch flags of the peer-memory transport.
    Every rank holds the whole batch too and checks each step against the single-GPU path of the same library."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), VIPANT_TRANSPORT="p2p", VIPANT_REQUIRE_P2P="1")
    torch.cuda.set_device(0 if torch.cuda.device_count() < world else rank)
    shared = torch.cuda.device_count() < world
    dist.init_process_group("gloo" if shared else "nccl", rank=rank, world_size=world)
    try:
        import vipant_b200 as vb
        b = B // world
        gen = torch.Generator(device="cuda").manual_seed(7)           # same stream of batches on every rank
        worst = 0.0
        for i in range(steps):
            f1 = torch.randn(B, D, device="cuda", generator=gen)
            f2 = 0.3 * f1 + 0.7 * torch.randn(B, D, device="cuda", generator=gen)
            ls_v = math.log(1 / 0.07) + 0.01 * i
            x1 = f1[rank * b:(rank + 1) * b].clone().requires_grad_(True)
            x2 = f2[rank * b:(rank + 1) * b].clone().requires_grad_(True)
            ls = torch.tensor(ls_v, device="cuda", requires_grad=True)
            loss = vb.infonce_loss(x1, x2, ls, group=dist.group.WORLD)
            loss.backward()
            g1, g2 = f1.clone().requires_grad_(True), f2.clone().requires_grad_(True)
            ls2 = torch.tensor(ls_v, device="cuda", requires_grad=True)
            full = vb.infonce_loss(g1, g2, ls2)
            full.backward()
            e = max(abs(loss.item() - full.item()) / abs(full.item()),
                    float((x1.grad - g1.grad[rank * b:(rank + 1) * b]).norm() / g1.grad[rank * b:(rank + 1) * b].norm()),
                    float((x2.grad - g2.grad[rank * b:(rank + 1) * b]).norm() / g2.grad[rank * b:(rank + 1) * b].norm()),
                    abs(ls.grad.item() - ls2.grad.item()) / abs(ls2.grad.item()))
            worst = max(worst, e)
        torch.cuda.synchronize()
        out[rank] = worst
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_p2p_fifty_steps_changing_data():
    """Sharded result == unsharded result of the same kernels on every one of 50 steps (fp32 summation order differs between
    the two decompositions: 1e-4)."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_many_steps, args=(world, _free_port(), 1024, 512, 50, out), nprocs=world, join=True)
    assert len(out) == world and max(out.values()) <= 1e-4, dict(out)


def _worker_composite(rank, world, port, out):
    """Three InfoNCE pairs per step under gather=True (VALCELossHead va + lv + al): every pair owns a peer-memory segment
    (one shared segment keeps two steps only and the first pair's backward failed), and ddp_average returns the
    per-rank partial d logit_scale."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), VIPANT_TRANSPORT="p2p", VIPANT_REQUIRE_P2P="1")
    torch.cuda.set_device(0 if torch.cuda.device_count() < world else rank)
    shared = torch.cuda.device_count() < world
    dist.init_process_group("gloo" if shared else "nccl", rank=rank, world_size=world)
    try:
        import vipant_b200 as vb
        from oracle.reference_loader import Cfg
        B, D = 512, 512
        b = B // world
        rng = np.random.default_rng(11)
        feats = [rng.standard_normal((B, D)).astype(np.float32) for _ in range(3)]
        res = {}
        for ddp in (False, True):
            head = vb.build_loss_head(Cfg(name="VALCELossHead", scaling=True, scale_max=None, va=True, lv=True, al=True,
                                          gather=True, ddp_average=ddp)).cuda().train()
            xs = [torch.from_numpy(f[rank * b:(rank + 1) * b]).cuda().requires_grad_(True) for f in feats]
            for _ in range(2):        # two steps: segments are reused
                for x in xs:
                    x.grad = None
                head.zero_grad()
                loss = head(*xs, normalized=False)
                loss.backward()
            torch.cuda.synchronize()
            res[ddp] = (loss.item(), [x.grad.cpu().numpy() for x in xs],
                        [getattr(head, a).logit_scale.grad.item() for a in ("loss_head_va", "loss_head_lv", "loss_head_al")])
        out[rank] = res
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_three_pairs_per_step_and_ddp_average():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_composite, args=(world, _free_port(), out), nprocs=world, join=True)
    B, D, b = 512, 512, 256
    rng = np.random.default_rng(11)
    feats = [rng.standard_normal((B, D)).astype(np.float32) for _ in range(3)]
    pairs = [(0, 1), (0, 2), (1, 2)]            # va, lv, al
    refs = [io.infonce_closed_form(feats[i], feats[j]) for i, j in pairs]
    want_loss = sum(r.loss for r in refs)
    grads = [np.zeros((B, D)) for _ in range(3)]
    for (i, j), r in zip(pairs, refs):
        grads[i] += r.dx1
        grads[j] += r.dx2

    def rel(a, r):
        return float(np.linalg.norm(a.astype(np.float64) - r) / np.linalg.norm(r))
    for rank in range(world):
        loss, g, dls = out[rank][False]
        assert abs(loss - want_loss) <= 1e-3 * want_loss
        for m in range(3):
            assert rel(g[m], grads[m][rank * b:(rank + 1) * b]) <= 1e-2
        for k, r in enumerate(refs):
            assert abs(dls[k] - r.dlogit_scale) <= 1e-2 * abs(r.dlogit_scale)
    # ddp_average: loss x world, gradients x world, d logit_scale = world x (per-rank partial): the MEAN over ranks (what DDP
    # computes) equals the single-process gradient for the features' encoders and for logit_scale alike
    for k, r in enumerate(refs):
        mean_dls = np.mean([out[rank][True][2][k] for rank in range(world)])
        assert abs(mean_dls - r.dlogit_scale) <= 1e-2 * abs(r.dlogit_scale)
    for rank in range(world):
        assert abs(out[rank][True][0] - world * want_loss) <= 1e-3 * world * want_loss
        assert rel(out[rank][True][1][0] / world, grads[0][rank * b:(rank + 1) * b]) <= 1e-2
