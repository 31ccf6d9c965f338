"""Packed embedding shards against the reference's on-disk cache: one np.savez_compressed(name, v=feat)
per clip (cvap/monitor/clap.py:54-61, audioset_clf.py:77-81), read back with np.load(f)["v"] (cvap/data/audioset_cls.py:337)."""
import os
import warnings

import numpy as np
import pytest
import torch

from vipant_b200 import embed_cache as ec


def _reference_cache(root, n, d=512, seed=0):
    """What the reference's encode_text / encode_audios leave on disk: per-clip npz, key "v", (k, D) or (D,) fp32."""
    rng = np.random.default_rng(seed)
    items = {}
    for i in range(n):
        k = int(rng.integers(1, 6))
        feat = rng.standard_normal((k, d)).astype(np.float32) if i % 3 else rng.standard_normal(d).astype(np.float32)
        name = f"clip_{i:05d}"
        np.savez_compressed(os.path.join(root, name), v=feat)
        items[name] = feat
    return items


def test_fp32_shard_is_bit_identical_to_the_npz_cache(tmp_path):
    root = tmp_path / "npz"
    root.mkdir()
    items = _reference_cache(str(root), 40)
    info = ec.pack_npz_dir(str(root), str(tmp_path / "a.vpae"))
    assert info["items"] == 40 and info["dim"] == 512
    shard = ec.EmbeddingShard(str(tmp_path / "a.vpae"))
    assert len(shard) == 40 and shard.names == sorted(items)
    for name, feat in items.items():
        ref = np.load(os.path.join(str(root), name + ".npz"))["v"]          # the reference's read
        got = shard[name]
        assert got.dtype == ref.dtype and got.shape == ref.shape and np.array_equal(got, ref)
        assert name in shard
    assert shard.rows.shape == (info["rows"], 512) and shard.rows.dtype == np.float32


def test_bf16_rounding_matches_torch():
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.standard_normal(100000).astype(np.float32) * 10.0 ** rng.integers(-30, 30, 100000),
                        np.array([0.0, -0.0, np.inf, -np.inf, 1.0, 1.00390625, 1.01171875, 3.3895314e38, 1e-40], np.float32)])
    want = torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    got = ec.to_bf16_bits(x)
    assert np.array_equal(got, want)                                       # round to nearest even, ties included
    assert np.array_equal(ec.from_bf16_bits(got), torch.from_numpy(x).to(torch.bfloat16).float().numpy())
    nan = ec.to_bf16_bits(np.array([np.nan], np.float32))
    assert np.isnan(ec.from_bf16_bits(nan))[0]


def test_bf16_shard_and_gather_order(tmp_path):
    root = tmp_path / "npz"
    root.mkdir()
    items = _reference_cache(str(root), 25, d=64, seed=3)
    ec.pack_npz_dir(str(root), str(tmp_path / "b.vpae"), dtype=ec.DTYPE_BF16)
    shard = ec.EmbeddingShard(str(tmp_path / "b.vpae"))
    assert shard.rows.dtype == np.uint16
    for name, feat in items.items():
        want = torch.from_numpy(np.ascontiguousarray(feat)).to(torch.bfloat16).float().numpy()
        assert np.array_equal(shard[name], want)
    # collator order: all rows of the first name, then the second, ... (cvap/data/audio_text.py:126-132)
    order = ["clip_00007", "clip_00001", "clip_00020", "clip_00007"]
    batch, counts = shard.gather(order)
    want = np.concatenate([np.atleast_2d(items[n]) for n in order])
    assert counts.tolist() == [np.atleast_2d(items[n]).shape[0] for n in order]
    assert np.array_equal(ec.from_bf16_bits(batch), torch.from_numpy(want).to(torch.bfloat16).float().numpy())
    out = np.zeros((64, 64), np.uint16)
    got, _ = shard.gather(order, out=out)
    assert got.base is out or got is out or np.shares_memory(got, out)
    with pytest.raises(ValueError):
        shard.gather(order, out=np.zeros((2, 64), np.uint16))


def test_missing_and_corrupt_items_follow_the_reference(tmp_path):
    root = tmp_path / "npz"
    root.mkdir()
    _reference_cache(str(root), 5, d=32)
    with open(os.path.join(str(root), "broken.npz"), "wb") as fw:
        fw.write(b"not a zip")
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        info = ec.pack_npz_dir(str(root), str(tmp_path / "c.vpae"))
    assert info["items"] == 5 and any("broken" in str(x.message) for x in w)
    shard = ec.EmbeddingShard(str(tmp_path / "c.vpae"))
    with pytest.warns(UserWarning, match="use random vector"):           # audioset_cls.py:339-341: random (D,) fp32 + warning
        v = shard.get("broken")
    assert v.shape == (32,) and v.dtype == np.float32 and 0.0 <= v.min() and v.max() < 1.0
    with pytest.raises(KeyError):
        shard["broken"]


def test_empty_shard_and_bad_files(tmp_path):
    info = ec.pack_items([], str(tmp_path / "e.vpae"))
    assert info["items"] == 0 and info["rows"] == 0
    shard = ec.EmbeddingShard(str(tmp_path / "e.vpae"))
    assert len(shard) == 0 and shard.rows.shape[0] == 0
    batch, counts = shard.gather([])
    assert batch.shape[0] == 0 and counts.size == 0
    with open(tmp_path / "junk.vpae", "wb") as fw:
        fw.write(b"\0" * 100)
    with pytest.raises(ValueError, match="not a vipant_b200 embedding shard"):
        ec.EmbeddingShard(str(tmp_path / "junk.vpae"))
    ec.pack_items([("a", np.ones((2, 8), np.float32))], str(tmp_path / "t.vpae"))
    raw = open(tmp_path / "t.vpae", "rb").read()
    with open(tmp_path / "t.vpae", "wb") as fw:
        fw.write(raw[:-8])
    with pytest.raises(ValueError, match="truncated"):
        ec.EmbeddingShard(str(tmp_path / "t.vpae"))
    with pytest.raises(ValueError, match="feature dim"):
        ec.pack_items([("a", np.ones(8, np.float32)), ("b", np.ones(9, np.float32))], str(tmp_path / "x.vpae"))
    with pytest.raises(ValueError, match="duplicate"):
        ec.pack_items([("a", np.ones(8, np.float32)), ("a", np.ones(8, np.float32))], str(tmp_path / "x.vpae"))
