"""Packed, memory-mappable shards of pre-computed embeddings.

The reference caches encoder outputs as ONE compressed npz per clip -- written by
``np.savez_compressed(f"{root}/{name}", v=feat)`` (``cvap/monitor/clap.py:54-61``, ``audioset_clf.py:77-81``) and read
back item by item with ``np.load(fname)["v"]`` (``cvap/data/audioset_cls.py:332-350``, ``audiocaps.py:117-123``), with a
random vector substituted when a file is missing or corrupt.  Every access inflates a zip member; a global batch of
32768 pairs is 65536 file opens.  A shard stores the same arrays back to back:

    [ header 64 B | item table n_items x 24 B | names blob | pad to 4096 | rows: (n_rows, D) fp32 or bf16, row-major ]

so that the payload is one ``np.memmap`` whose slices feed ``vpa_normalize_cast`` (fp32 or bf16 input) without a copy on
the host side other than the pinned staging buffer.  ``EmbeddingShard[name]`` returns exactly what ``np.load(f)["v"]``
returns for fp32 shards (same shape, dtype, bits); bf16 shards hold the round-to-nearest-even bf16 of those values (what
``torch.Tensor.to(torch.bfloat16)`` produces) and halve the bytes -- the tensor-core path rounds its operands to bf16 anyway
(after normalisation, so bf16 shards are for ALREADY NORMALISED features or for eval-side scoring where 3 significant
digits suffice; fp32 is the default).

Host-side only: no device code here.  ``gather(names)`` assembles a batch in the collators' order (all rows of item 0,
then item 1, ... -- ``cvap/data/audio_text.py:126-132``) into a caller-provided (pinned) buffer.
"""
from __future__ import annotations

import os
import struct
import warnings
import zipfile
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

MAGIC = b"VPAE"
VERSION = 1
DTYPE_F32, DTYPE_BF16 = 0, 1
_HEADER = struct.Struct("<4sIIIQQQQ16x")          # magic, version, dtype, D, n_items, n_rows, names_bytes, payload_offset
_ITEM = struct.Struct("<QQII")                     # name offset, first row, n_rows, n_dims (1: stored as a vector, 2: matrix)
_ALIGN = 4096

__all__ = ["EmbeddingShard", "pack_items", "pack_npz_dir", "to_bf16_bits", "from_bf16_bits", "DTYPE_F32", "DTYPE_BF16"]


def to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """fp32 -> bf16 bit patterns (uint16), round to nearest even; NaN stays NaN (quiet bit forced)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    rounded = (u + (np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1)))) >> np.uint32(16)
    nan = (u & np.uint32(0x7FFFFFFF)) > np.uint32(0x7F800000)
    return np.where(nan, (u >> np.uint32(16)) | np.uint32(0x0040), rounded).astype(np.uint16)


def from_bf16_bits(b: np.ndarray) -> np.ndarray:
    return (np.asarray(b, dtype=np.uint16).astype(np.uint32) << np.uint32(16)).view(np.float32)


def pack_items(items: Iterable[Tuple[str, np.ndarray]], path: str, dtype: int = DTYPE_F32) -> Dict[str, int]:
    """Write a shard from (name, array) pairs; arrays are (D,) or (k, D) as the reference's encoders produce them.
    Returns {"items", "rows", "dim", "bytes"}.  Item order is preserved (it is the row order of the payload)."""
    if dtype not in (DTYPE_F32, DTYPE_BF16):
        raise ValueError("dtype must be DTYPE_F32 or DTYPE_BF16")
    table: List[Tuple[int, int, int, int]] = []
    names = bytearray()
    rows: List[np.ndarray] = []
    n_rows, dim = 0, None
    seen = set()
    for name, arr in items:
        a = np.asarray(arr)
        if a.ndim not in (1, 2):
            raise ValueError(f"item {name!r}: expected a (D,) or (k, D) array, got shape {a.shape}")
        a2 = a.reshape(1, -1) if a.ndim == 1 else a
        if dim is None:
            dim = int(a2.shape[1])
        elif a2.shape[1] != dim:
            raise ValueError(f"item {name!r}: feature dim {a2.shape[1]} != {dim}")
        if name in seen:
            raise ValueError(f"duplicate item name {name!r}")
        seen.add(name)
        raw = name.encode("utf-8")
        table.append((len(names), n_rows, a2.shape[0], a.ndim))
        names += raw + b"\0"
        rows.append(np.ascontiguousarray(a2, dtype=np.float32))
        n_rows += a2.shape[0]
    dim = dim or 0
    head = _HEADER.size + _ITEM.size * len(table) + len(names)
    payload_offset = (head + _ALIGN - 1) // _ALIGN * _ALIGN
    tmp = path + ".tmp"
    with open(tmp, "wb") as fw:
        fw.write(_HEADER.pack(MAGIC, VERSION, dtype, dim, len(table), n_rows, len(names), payload_offset))
        for t in table:
            fw.write(_ITEM.pack(*t))
        fw.write(bytes(names))
        fw.write(b"\0" * (payload_offset - head))
        for a in rows:
            fw.write((to_bf16_bits(a) if dtype == DTYPE_BF16 else a).tobytes())
    os.replace(tmp, path)
    return {"items": len(table), "rows": n_rows, "dim": dim, "bytes": os.path.getsize(path)}


def pack_npz_dir(root: str, path: str, dtype: int = DTYPE_F32, names: Optional[Sequence[str]] = None,
                 key: str = "v") -> Dict[str, int]:
    """Convert a directory of the reference's per-clip ``{name}.npz`` files (array under ``key``) into one shard.
    Unreadable files are skipped with a warning (the reference substitutes a random vector at read time; a shard never
    stores made-up data -- ``EmbeddingShard.get`` reproduces the substitution for names it does not hold)."""
    if names is None:
        names = sorted(f[:-4] for f in os.listdir(root) if f.endswith(".npz"))

    def gen():
        for name in names:
            try:
                with np.load(os.path.join(root, name + ".npz")) as z:
                    yield name, z[key]
            except (OSError, KeyError, ValueError, zipfile.BadZipFile) as exc:
                warnings.warn(f"skipping {name}: {exc}")
    return pack_items(gen(), path, dtype)


class EmbeddingShard:
    """Read side: ``shard[name]`` == ``np.load(f"{root}/{name}.npz")["v"]`` for fp32 shards."""

    def __init__(self, path: str):
        self.path = path
        with open(path, "rb") as fr:
            head = fr.read(_HEADER.size)
            if len(head) < _HEADER.size:
                raise ValueError(f"{path}: truncated header")
            magic, version, dtype, dim, n_items, n_rows, names_bytes, payload_offset = _HEADER.unpack(head)
            if magic != MAGIC or version != VERSION:
                raise ValueError(f"{path}: not a vipant_b200 embedding shard (magic {magic!r}, version {version})")
            if dtype not in (DTYPE_F32, DTYPE_BF16):
                raise ValueError(f"{path}: unknown dtype code {dtype}")
            table = fr.read(_ITEM.size * n_items)
            blob = fr.read(names_bytes)
        es = 4 if dtype == DTYPE_F32 else 2
        if len(table) != _ITEM.size * n_items or len(blob) != names_bytes or \
                os.path.getsize(path) < payload_offset + n_rows * dim * es:
            raise ValueError(f"{path}: truncated shard")
        self.dtype, self.dim, self.n_rows = dtype, dim, n_rows
        self._index: Dict[str, Tuple[int, int, int]] = {}
        self.names: List[str] = []
        for i in range(n_items):
            off, first, cnt, nd = _ITEM.unpack_from(table, i * _ITEM.size)
            name = blob[off:blob.index(b"\0", off)].decode("utf-8")
            self._index[name] = (first, cnt, nd)
            self.names.append(name)
        np_dtype = np.float32 if dtype == DTYPE_F32 else np.uint16
        self.rows = np.memmap(path, dtype=np_dtype, mode="r", offset=payload_offset, shape=(n_rows, dim)) \
            if n_rows else np.zeros((0, dim), np_dtype)

    def __len__(self):
        return len(self.names)

    def __contains__(self, name):
        return name in self._index

    def row_range(self, name: str) -> Tuple[int, int]:
        first, cnt, _ = self._index[name]
        return first, first + cnt

    def __getitem__(self, name: str) -> np.ndarray:
        first, cnt, nd = self._index[name]
        block = self.rows[first:first + cnt]
        out = np.array(block) if self.dtype == DTYPE_F32 else from_bf16_bits(block)
        return out[0] if nd == 1 else out

    def get(self, name: str, rng: Optional[np.random.Generator] = None) -> np.ndarray:
        """``shard[name]``, or -- like the reference's readers (``audioset_cls.py:337-341``) -- a random (D,) fp32 vector
        with a warning when the shard does not hold `name`."""
        if name in self._index:
            return self[name]
        warnings.warn(f"use random vector instead because `{name}` is not in {self.path}.")
        r = rng.random(self.dim) if rng is not None else np.random.rand(self.dim)
        return r.astype("float32")

    def gather(self, names: Sequence[str], out: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray]:
        """Rows of `names` in order, concatenated (collator order): returns (batch, counts).  `out` may be a caller-owned
        (e.g. pinned) array of the shard's storage dtype (fp32, or uint16 bf16 bits) with at least the needed rows."""
        if len(names) == 0:
            return (np.empty((0, self.dim), dtype=self.rows.dtype) if out is None else out[:0]), np.zeros(0, np.int64)
        spans = np.asarray([self._index[n][:2] for n in names], dtype=np.int64)
        firsts, counts = spans[:, 0], spans[:, 1]
        total = int(counts.sum())
        if out is None:
            out = np.empty((total, self.dim), dtype=self.rows.dtype)
        elif out.dtype != self.rows.dtype or out.ndim != 2 or out.shape[1] != self.dim or out.shape[0] < total:
            raise ValueError("gather: `out` must be a 2-d array of the shard's storage dtype with enough rows")
        # row index of every output row: one vectorised take instead of a Python loop over the items
        starts = np.cumsum(counts) - counts
        idx = np.repeat(firsts - starts, counts) + np.arange(total, dtype=np.int64)
        np.take(self.rows, idx, axis=0, out=out[:total])
        return out[:total], counts
