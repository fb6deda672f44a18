"""The C-ABI library builds, loads and exports every symbol declared in include/panib200.h.

No device compute here (CPU container): only the pure-host helpers are called.
"""

from __future__ import annotations

import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

import __graft_entry__ as entry
from oracle import oracle

HEADER = entry.ROOT / "include" / "panib200.h"


@pytest.fixture(scope="module")
def lib() -> ctypes.CDLL:
    entry.build()
    from pyani_plus_b200 import engine

    return engine.load_library()


def declared_symbols() -> list[str]:
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(panib_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib: ctypes.CDLL) -> None:
    names = declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"{name} declared in panib200.h but not exported"


def test_header_cites_reference() -> None:
    text = HEADER.read_text()
    assert "pyani_plus/methods/sourmash.py:67-83" in text
    assert "pyani_plus/methods/sourmash.py:184-200" in text


def test_max_hash_matches_oracle_and_fixtures(lib: ctypes.CDLL) -> None:
    from pyani_plus_b200 import engine

    assert engine.max_hash(300) == 61489146912365176
    assert engine.max_hash(1000) == 18446744073709552
    for scaled in (1, 2, 3, 50, 100, 299, 300, 1000, 12345, 10**9, 2**40 + 1):
        assert engine.max_hash(scaled) == oracle.max_hash(scaled) == oracle.py_max_hash(scaled)


def test_plan_buckets_is_monotone_and_in_range(lib: ctypes.CDLL) -> None:
    from pyani_plus_b200 import engine

    for n_kmers, scaled in ((5_000_000, 1000), (5_000_000, 100), (40_000, 300), (18_000, 1), (100, 10**12)):
        nb, bmul = engine.plan_buckets(n_kmers, scaled)
        assert nb >= 1
        mh = engine.max_hash(scaled)
        for h in (1, mh // 3, mh // 2, mh - 1, mh):
            assert (h * bmul) >> 64 < nb
        assert (mh * bmul) >> 64 >= nb - 2  # the top bucket is used
    with pytest.raises(Exception, match="bad arguments"):
        engine.plan_buckets(-1, 1000)


def test_ani_host_matches_oracle_exactly(lib: ctypes.CDLL) -> None:
    from pyani_plus_b200 import engine

    rng = np.random.default_rng(1)
    qc = rng.integers(0, 6000, 40).astype(np.int32)
    sc = rng.integers(0, 6000, 30).astype(np.int32)
    ov = np.minimum(rng.integers(0, 6000, (40, 30)), np.minimum.outer(qc, sc)).astype(np.uint32)
    ov[3, :] = 0
    for k in (31, 21):
        ident, cov = engine.ani_host(ov, qc, sc, k)
        for i in range(40):
            for j in range(30):
                row = oracle.pair_row(int(ov[i, j]), int(qc[i]), int(sc[j]), k)
                if row is None:
                    assert np.isnan(ident[i, j]) and np.isnan(cov[i, j])
                else:
                    assert ident[i, j] == row["max_containment_ani"]
                    assert cov[i, j] == row["query_containment_ani"]


def test_engine_fails_loudly_without_gpu() -> None:
    """No CPU fallback: constructing the engine on a box without CUDA must raise."""
    import torch

    from pyani_plus_b200 import engine

    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(engine.EngineError, match="no CPU fallback"):
        engine.Engine()


def test_product_never_imports_oracle() -> None:
    """The oracle is test infrastructure; nothing under pyani_plus_b200/ may reference it."""
    for path in Path(entry.PKG).rglob("*"):
        if path.suffix in {".py", ".cu", ".cuh", ".cpp", ".h"}:
            text = path.read_text()
            assert "import oracle" not in text and "from oracle" not in text, path
            assert "liboracle" not in text, path


def _prototypes() -> dict[str, list[str]]:
    """name -> list of C parameter declarations, parsed from the header."""
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    out = {}
    for m in re.finditer(r"PANIB_API\s+[\w\s\*]+?\b(panib_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        params = [p.strip() for p in m.group(2).replace("\n", " ").split(",")]
        out[m.group(1)] = [] if params == ["void"] else params
    return out


def test_ctypes_binding_matches_header(lib: ctypes.CDLL) -> None:
    """Every bound function has as many argtypes as the header has parameters, pointers are bound as
    pointers and 64-bit integers as 64-bit (a drifted binding would pass truncated pointers)."""
    protos = _prototypes()
    assert len(protos) >= 20
    checked = 0
    for name, params in protos.items():
        fn = getattr(lib, name)
        if fn.argtypes is None:
            assert not params or name in ("panib_version", "panib_device_count", "panib_launch_count"), name
            continue
        assert len(fn.argtypes) == len(params), f"{name}: {len(fn.argtypes)} argtypes, {len(params)} parameters"
        for ctype, decl in zip(fn.argtypes, params, strict=True):
            is_ptr = "*" in decl
            size = ctypes.sizeof(ctype)
            if is_ptr or re.search(r"\b(u?int64_t|size_t|double)\b", decl):
                assert size == 8, f"{name}: '{decl}' bound as {ctype}"
            else:
                assert size == 4, f"{name}: '{decl}' bound as {ctype}"
        checked += 1
    assert checked >= 15


def test_pack_host_code_paths_agree(lib: ctypes.CDLL) -> None:
    """panib_pack_host: scalar, AVX2, AVX-512 and the threaded pool give the words of the device pack
    (pack.cuh, here through the host emulation library) -- lower case, N runs, non-letters included."""
    from pyani_plus_b200 import engine

    emu = ctypes.CDLL(str(entry.PKG / "libpanib_hostemu.so"))
    emu.emu_pack_ascii.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
    rng = np.random.default_rng(7)
    n = 3 * 262144 + 64 * 1021  # several pool blocks plus a ragged tail
    alphabet = np.frombuffer(b"ACGTacgtNnRY-\n*", dtype=np.uint8)
    a = rng.choice(alphabet, size=n, p=[.2, .2, .2, .2, .04, .04, .04, .04, .01, .01, .004, .004, .004, .004, .004])
    a[5000:5040] = ord("N")
    want_p = np.zeros(n // 16, np.uint32)
    want_m = np.zeros(n // 32, np.uint32)
    emu.emu_pack_ascii(a.ctypes.data, n, want_p.ctypes.data, want_m.ctypes.data)
    ran = 0
    for threads in (-1, -2, -3, 0, 1, 3):
        try:
            got_p, got_m = engine.pack_host(a, threads)
        except engine.EngineError as exc:  # an instruction set this CPU lacks
            assert "lacks" in str(exc)
            continue
        assert (got_p == want_p).all() and (got_m == want_m).all(), threads
        ran += 1
    assert ran >= 4
    assert lib.panib_host_threads() >= 1
    with pytest.raises(ValueError, match="multiple of 32"):
        engine.pack_host(a[:33])


def test_pack_host_tiles_sparse_mask(lib: ctypes.CDLL) -> None:
    """panib_pack_host_tiles (the form the ingest pipeline runs): the packed words of panib_pack_host, a dirty
    flag for exactly the tiles that hold an invalid base, their 128 mask words written and a clean tile's
    mask words left untouched -- unaligned buffers (no non-temporal stores) included."""
    from pyani_plus_b200 import engine

    engine.load_library()
    tile = 4096
    n_tiles = 64 * 3 + 17  # several pool blocks plus a ragged one
    rng = np.random.default_rng(11)
    a = rng.choice(np.frombuffer(b"ACGTacgt", dtype=np.uint8), size=n_tiles * tile)
    dirty_tiles = sorted({0, 5, 63, 64, 100, n_tiles - 1})
    for t in dirty_tiles:
        at = t * tile + int(rng.integers(0, tile - 40))
        a[at: at + int(rng.integers(1, 40))] = ord("N")
    a[5 * tile: 6 * tile] = ord("n")  # a whole tile of invalid bases
    want_p, want_m = engine.pack_host(a, 0)
    for threads, shift in ((0, 0), (1, 0), (3, 1)):
        pbuf = np.zeros(a.size // 16 + 4, np.uint32)
        mbuf = np.full(a.size // 32 + 4, 0xABABABAB, np.uint32)
        got_p, got_m = pbuf[shift: shift + a.size // 16], mbuf[shift: shift + a.size // 32]
        dirty = np.full(n_tiles, 7, np.uint8)
        rc = lib.panib_pack_host_tiles(a.ctypes.data, a.size, got_p.ctypes.data, got_m.ctypes.data, dirty.ctypes.data,
                                       threads)
        assert rc == 0
        assert (got_p == want_p).all()
        assert np.flatnonzero(dirty).tolist() == dirty_tiles and set(dirty.tolist()) <= {0, 1}
        gm, wm = got_m.reshape(n_tiles, tile // 32), want_m.reshape(n_tiles, tile // 32)
        for t in range(n_tiles):
            if dirty[t]:
                assert (gm[t] == wm[t]).all(), t
            else:
                assert (gm[t] == 0xABABABAB).all() and not wm[t].any(), t
    assert lib.panib_pack_host_tiles(a.ctypes.data, tile + 32, None, None, None, 0) != 0
    assert lib.panib_ingest_scratch_bytes(0) == 0 and lib.panib_ingest_scratch_bytes(64 * tile * 1000) > 0
