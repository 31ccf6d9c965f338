"""The C-ABI library loads on a CPU-only machine and exports every symbol include/vipant_b200.h declares.

No compute call is made here (there is no GPU); argument validation paths that return before touching the
device ARE exercised, because they are part of the drop-in contract (errors are codes, never exceptions/aborts).
"""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from vipant_b200 import _cabi, build

HEADER = os.path.join(ROOT, "include", "vipant_b200.h")


@pytest.fixture(scope="module")
def lib():
    build.build()          # no-op when up to date; nvcc cross-compiles sm_100a without a GPU
    return _cabi.lib()


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vpa_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_cabi.EXPORTED_SYMBOLS)


def test_every_declared_symbol_is_exported(lib):
    raw = ctypes.CDLL(_cabi.library_path())
    for name in declared_symbols():
        assert hasattr(raw, name), f"{name} declared in include/vipant_b200.h but not exported"


def test_version(lib):
    assert lib.vpa_version() == 100


def test_only_c_symbols_are_public():
    """The boundary is extern "C": no torch / C++-mangled API symbols are exported by name vpa_*."""
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", _cabi.library_path()], capture_output=True, text=True).stdout
    names = [l.split()[-1] for l in out.splitlines() if l.strip()]
    assert all(n in names for n in declared_symbols())
    assert not any("torch" in n or "c10" in n for n in names)


def test_workspace_queries_are_pure(lib):
    assert lib.vpa_infonce_workspace_bytes(0, 0, 512, 0) == 0
    assert lib.vpa_infonce_workspace_bytes(512, 512, 512, 0) > 0
    assert lib.vpa_infonce_workspace_bytes(512, 512, 512, 1) > 0
    assert lib.vpa_sim_workspace_bytes(975, 4875) >= 975 * 4875 * 4
    assert lib.vpa_sim_workspace_bytes(0, 10) == 0
    assert lib.vpa_infonce_host_scratch_bytes(64, 512, 0) > 2 * 64 * 512 * 4


def test_invalid_arguments_return_codes_not_crashes(lib):
    # null pointers / bad shapes are rejected before any device work
    assert lib.vpa_normalize_cast(None, 0, 4, 512, 512, 0, None, None, None, None) == -1
    assert b"null" in lib.vpa_last_error_string()
    one = ctypes.c_float(0.0)
    p = ctypes.addressof(one)
    assert lib.vpa_normalize_cast(p, 0, 4, 510, 510, 0, None, None, None, None) == -1       # D % 4
    assert lib.vpa_normalize_cast(p, 7, 4, 512, 512, 0, None, None, None, None) == -1       # dtype
    assert lib.vpa_infonce_loss(None, None, None, 8, None, None) == -1
    # D = 96 is not covered by the tensor-core tiling -> UNSUPPORTED (-3), fp32 path accepts it
    rc = lib.vpa_infonce_fwd(p, p, p, p, 0, 8, 8, 96, 0, p, 0.0, p, p, 1 << 20, p, p, p, p, None)
    assert rc == -3
    rc = lib.vpa_infonce_fwd(p, p, p, p, 0, 8, 4, 512, 0, p, 0.0, p, p, 1 << 20, p, p, p, p, None)
    assert rc == -1                                                                          # rows_global < rows_local
    rc = lib.vpa_infonce_fwd(p, p, p, p, 0, 512, 512, 512, 0, p, 0.0, p, p, 16, p, p, p, p, None)
    assert rc == -2                                                                          # workspace too small
    assert lib.vpa_sim_rank_topk(p, p, 4, 4, 512, 512, 512, None, 0, 1, None, None, None, p, 0, None) == -2


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_cabi, "_LIB", None)
    monkeypatch.setattr(build, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_cabi.VipantB200Error, match="no CPU or PyTorch fallback"):
        _cabi.lib()


@pytest.mark.parametrize("b,B", [(32768, 32768), (16384, 32768), (8192, 32768), (4096, 32768), (512, 512), (640, 640),
                                 (384, 1152), (64, 64), (1000, 1000), (4096, 4096), (2048, 16384), (300, 2400)])
@pytest.mark.parametrize("peer_memory", [0, 1])
def test_sweep_plan_covers_every_tile(b, B, peer_memory):
    """vpa_plan_query (host only): every row block's tiles are covered exactly once by n_big equal chunks + one short tail,
    also in the plan of the peer-memory transport (relay CTAs take SM pairs from the sweeps)."""
    import ctypes
    from vipant_b200 import _cabi
    lib = _cabi.lib()
    out = (ctypes.c_int * 10)()
    assert lib.vpa_plan_query(b, B, 512, _cabi.PREC_BF16_TC, peer_memory, out) == 0
    n_tiles, f_chunks, f_tpc, f_small, b_chunks, b_tpc, b_small, f_iblk, b_iblk, impl = list(out)
    assert impl == 1 and n_tiles == -(-B // 256) and f_iblk == -(-b // 256) and b_iblk == -(-b // 128)
    for chunks, tpc, small in ((f_chunks, f_tpc, f_small), (b_chunks, b_tpc, b_small)):
        n_big = chunks - (1 if small else 0)
        big_tiles = n_tiles - small
        assert n_big >= 1 and 0 <= small < max(tpc, 1) + (small == 0)
        assert (n_big - 1) * tpc < big_tiles <= n_big * tpc          # no empty chunk, nothing left over


def test_peer_memory_entry_points_reject_bad_arguments(lib):
    """The vpa_p2p_* surface validates before it touches the device: world range, null handles, stale steps."""
    handle = ctypes.c_void_p()
    ipc = ctypes.create_string_buffer(64)
    assert lib.vpa_p2p_create(512, 1, 0, 512, 0, ctypes.byref(handle), ipc) == -1        # needs >= 2 ranks
    assert b"world" in lib.vpa_last_error_string()
    assert lib.vpa_p2p_create(512, 9, 0, 512, 0, ctypes.byref(handle), ipc) == -1        # one NVSwitch node: <= 8
    assert lib.vpa_p2p_create(512, 4, 4, 512, 0, ctypes.byref(handle), ipc) == -1        # rank out of range
    assert lib.vpa_p2p_create(512, 2, 0, 96, 0, ctypes.byref(handle), ipc) == -3         # D not covered by the tensor-core tiling
    assert lib.vpa_p2p_connect(None, ipc) == -1
    assert lib.vpa_p2p_destroy(None) == 0
    one = ctypes.c_float(0.0)
    p = ctypes.addressof(one)
    epoch = ctypes.c_uint32()
    assert lib.vpa_infonce_fwd_p2p(None, p, p, 0, 512, 2, 0, 512, 512, 512, 0, p, 0.0, 0, p, ctypes.byref(epoch), None) == -1
    assert lib.vpa_infonce_bwd_p2p(None, 1, p, p, 0, 512, 2, 0, 512, 512, 512, 0, 0, p, p, p, p, 1, None) == -1


def test_transport_selection(monkeypatch):
    """functional._transport: one GPU -> local; a process group of <= 8 ranks -> peer memory unless overridden."""
    from vipant_b200 import functional as F_

    class FakeDist:
        def __init__(self, world, backend):
            self.world, self.backend = world, backend

        def get_world_size(self, group=None):
            return self.world

        def get_backend(self, group=None):
            return self.backend
    monkeypatch.delenv("VIPANT_TRANSPORT", raising=False)
    monkeypatch.delenv("VIPANT_HOST_ORCHESTRATION", raising=False)
    assert F_._transport(None) == "local"
    monkeypatch.setattr(F_, "dist", FakeDist(1, "nccl"))
    assert F_._transport(object()) == "local"
    monkeypatch.setattr(F_, "dist", FakeDist(8, "nccl"))
    assert F_._transport(object()) == "p2p"
    monkeypatch.setattr(F_, "dist", FakeDist(16, "nccl"))
    assert F_._transport(object()) == "nccl"
    monkeypatch.setattr(F_, "dist", FakeDist(16, "gloo"))
    assert F_._transport(object()) == "host"
    monkeypatch.setattr(F_, "dist", FakeDist(4, "nccl"))
    monkeypatch.setenv("VIPANT_TRANSPORT", "nccl")
    assert F_._transport(object()) == "nccl"
    monkeypatch.setenv("VIPANT_HOST_ORCHESTRATION", "1")
    assert F_._transport(object()) == "host"


@pytest.mark.parametrize("source_major", [0, 1])
@pytest.mark.parametrize("world,b", [(8, 4096), (4, 8192), (2, 16384), (3, 384), (2, 300), (4, 512), (2, 64), (5, 1000)])
def test_relay_items_cover_every_peer_row_once(lib, world, b, source_major):
    """The relay CTAs' item map (the compiled relay_item_decode, evaluated on the host): every row of every peer block is
    pulled exactly once per matrix, one item per 256-row chunk (what its arrival flag stands for).  Chunk-major (forward):
    the x2 operands come first and chunk k of all peers before chunk k+1 of any.  Source-major (backward, x1 operands only):
    peer me+1 completely, then me+2, ...  Out-of-range items are rejected."""
    cpr = -(-b // 256)
    out = (ctypes.c_int * 5)()
    for me in (0, world - 1):
        for m0 in (0, 1):
            total = (2 - m0) * cpr * (world - 1)
            seen = np.zeros((2, world, b), dtype=np.int32)
            chunks = set()
            order = []
            for item in range(total):
                assert lib.vpa_debug_relay_item(item, m0, source_major, world, me, cpr, b, out) == 0
                m, src, c, row0, rows = list(out)
                assert m0 <= m < 2 and 0 <= src < world and src != me and 0 <= c < cpr and rows >= 1
                assert row0 == c * 256 and row0 + rows == min(b, (c + 1) * 256)
                seen[m, src, row0:row0 + rows] += 1
                chunks.add((m, src, c))
                order.append((m, (src - me) % world, c) if source_major else (m, c))
            peers = [q for q in range(world) if q != me]
            assert np.all(seen[m0:, peers, :] == 1) and np.all(seen[:, me, :] == 0) and np.all(seen[:m0] == 0)
            assert len(chunks) == total
            assert order == sorted(order)             # matrix-major, then chunk-major / peer-major starting at me+1
            assert lib.vpa_debug_relay_item(total, m0, source_major, world, me, cpr, b, out) == -1
