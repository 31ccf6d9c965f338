"""Embedding-cache shards (vipant_b200/embed_cache.py) feeding the device path: rows gathered from a
shard into a PINNED staging buffer go to the GPU as they are stored (fp32, or bf16 bits viewed as torch.bfloat16) and
straight into the kernels -- the training loss and the retrieval scoring -- without a host-side conversion."""
import numpy as np
import pytest
import torch

from oracle import infonce_oracle as io
from oracle import retrieval_oracle as ro

pytestmark = pytest.mark.gpu


def _shards(tmp_path, dtype_code, n=384, D=512):
    from vipant_b200 import embed_cache as ec
    rng = np.random.default_rng(5)
    a = rng.standard_normal((n, D)).astype(np.float32)
    t = (0.5 * a.repeat(5, axis=0) + rng.standard_normal((5 * n, D))).astype(np.float32)
    names = [f"clip{i:05d}" for i in range(n)]
    pa, pt = str(tmp_path / "audio.vpae"), str(tmp_path / "text.vpae")
    ec.pack_items(((names[i], a[i]) for i in range(n)), pa, dtype=dtype_code)              # one vector per clip
    ec.pack_items(((names[i], t[5 * i:5 * i + 5]) for i in range(n)), pt, dtype=dtype_code)  # five captions per clip
    return ec.EmbeddingShard(pa), ec.EmbeddingShard(pt), names, a, t


def _to_device(shard, names):
    from vipant_b200 import embed_cache as ec
    total = sum(shard.row_range(n)[1] - shard.row_range(n)[0] for n in names)
    stage = torch.empty((total, shard.dim), dtype=torch.float32 if shard.dtype == ec.DTYPE_F32 else torch.uint16).pin_memory()
    rows, counts = shard.gather(names, out=stage.numpy())
    assert rows.shape[0] == total and counts.sum() == total
    dev = stage.to("cuda", non_blocking=True)
    return dev if shard.dtype == ec.DTYPE_F32 else dev.view(torch.bfloat16)


@pytest.mark.parametrize("code", ["f32", "bf16"])
def test_shard_rows_feed_training_loss_and_scoring(tmp_path, code):
    import vipant_b200 as vb
    from vipant_b200 import embed_cache as ec
    from vipant_b200 import functional as F_
    dtype_code = ec.DTYPE_F32 if code == "f32" else ec.DTYPE_BF16
    sa, st, names, a, t = _shards(tmp_path, dtype_code)
    order = list(np.random.default_rng(1).permutation(len(names)))          # a shuffled batch, collator order
    batch = [names[i] for i in order]
    xa = _to_device(sa, batch)
    xt_all = _to_device(st, batch)                                           # (5 n, D): captions of clip 0, clip 1, ...
    # what the shard holds (bf16 shards: the round-to-nearest-even bf16 of the features)
    ha = np.stack([sa[n] for n in batch]).astype(np.float32)
    ht = np.concatenate([st[n] for n in batch]).astype(np.float32)
    assert np.array_equal(xa.float().cpu().numpy(), ha) and np.array_equal(xt_all.float().cpu().numpy(), ht)
    # ---- training step on (audio, first caption) pairs, inputs in the shard's storage dtype
    x1 = xa.clone().requires_grad_(True)
    x2 = xt_all[0::5].clone().requires_grad_(True)
    ls = torch.tensor(float(np.log(1 / 0.07)), device="cuda", requires_grad=True)
    loss = vb.infonce_loss(x1, x2, ls, precision="bf16")
    loss.backward()
    ref = io.infonce_closed_form(ha, ht[0::5])
    assert abs(loss.item() - ref.loss) <= 1e-3 * abs(ref.loss)
    tol = 1e-2 if code == "f32" else 2e-2             # bf16 gradients are themselves rounded to 8 bits
    assert np.linalg.norm(x1.grad.float().cpu().numpy() - ref.dx1) <= tol * np.linalg.norm(ref.dx1)
    assert x1.grad.dtype == x1.dtype
    # ---- retrieval scoring 1-vs-5 on the same rows: ranks against the fp64 oracle on the stored values
    an, tn = vb.l2_normalize(xa), vb.l2_normalize(xt_all)
    n = len(batch)
    res = F_.sim_rank_fused(an, tn, gt_q=torch.arange(5 * n, device="cuda").view(n, 5), gt_k=torch.arange(5 * n, device="cuda") // 5)
    S = ro.normalize(ha.astype(np.float64)) @ ro.normalize(ht.astype(np.float64)).T
    gt12 = np.arange(5 * n).reshape(n, 5)
    want12, want21 = ro.rank_of(S, gt12), ro.rank_of(S.T, (np.arange(5 * n) // 5)[:, None])
    got12, got21 = res["ranks_q"].cpu().numpy(), res["ranks_k"].cpu().numpy()
    assert (got12 != want12).mean() < 2e-3 and (got21 != want21).mean() < 2e-3 and np.abs(got12 - want12).max() <= 1
