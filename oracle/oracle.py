"""CPU ORACLE for the pyani-plus sourmash path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(``pyani_plus_b200``) must never import it; it has no CPU fallback.

Two independent restatements live here so that each can pin the other:

* ``liboracle.so`` (``panib_oracle.c``; byte strings + memcmp + qsort) via ctypes, and
* a pure-Python murmur/sketch (``py_*`` functions; big-int arithmetic) for tiny inputs.

Both follow the conventions listed in SURVEY.md section 8c, which restate the behaviour of
``sourmash scripts singlesketch`` / ``manysearch`` as called from
``pyani_plus/methods/sourmash.py:67-83,184-200`` of the reference, and both are pinned against
the reference's own fixture files by ``tests/test_oracle_golden.py``.
"""

from __future__ import annotations

import ctypes
import gzip
import hashlib
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "liboracle.so"
_lib = None

c_u64p = ctypes.POINTER(ctypes.c_uint64)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_u8p = ctypes.POINTER(ctypes.c_uint8)


def build(force: bool = False) -> Path:
    """Compile liboracle.so with the committed Makefile (gcc, OpenMP)."""
    src = _HERE / "panib_oracle.c"
    if force or not _LIB_PATH.is_file() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-C", str(_HERE), "-B"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    """Load (building if needed) the C oracle."""
    global _lib  # noqa: PLW0603
    if _lib is None:
        build()
        L = ctypes.CDLL(str(_LIB_PATH))
        L.oracle_max_hash.restype = ctypes.c_uint64
        L.oracle_max_hash.argtypes = [ctypes.c_uint64]
        L.oracle_murmur64.restype = ctypes.c_uint64
        L.oracle_murmur64.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_uint32]
        L.oracle_sketch.restype = ctypes.c_int64
        L.oracle_sketch.argtypes = [
            ctypes.c_void_p, c_i64p, ctypes.c_int64, ctypes.c_int, ctypes.c_uint64,
            ctypes.c_uint32, c_u64p, ctypes.c_int64,
        ]
        L.oracle_sketch_fast.restype = ctypes.c_int64
        L.oracle_sketch_fast.argtypes = L.oracle_sketch.argtypes
        L.oracle_intersect.restype = ctypes.c_int64
        L.oracle_intersect.argtypes = [c_u64p, ctypes.c_int64, c_u64p, ctypes.c_int64]
        L.oracle_ani_from_containment.restype = ctypes.c_double
        L.oracle_ani_from_containment.argtypes = [ctypes.c_double, ctypes.c_int]
        L.oracle_pair_row.restype = ctypes.c_int
        L.oracle_pair_row.argtypes = [
            ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
            ctypes.POINTER(ctypes.c_double),
        ]
        L.oracle_synth_threshold.restype = ctypes.c_uint64
        L.oracle_synth_threshold.argtypes = [ctypes.c_uint64, ctypes.c_uint64]
        L.oracle_synth_genome.restype = None
        L.oracle_synth_genome.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int64, ctypes.c_void_p]
        L.oracle_num_threads.restype = ctypes.c_int
        L.oracle_synth_sketch_batch.restype = ctypes.c_int64
        L.oracle_synth_sketch_batch.argtypes = [
            ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
            ctypes.c_uint64, c_u64p, ctypes.c_int64, c_i64p,
        ]
        L.oracle_sketch_batch.restype = None
        L.oracle_sketch_batch.argtypes = [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_uint64,
            c_u64p, ctypes.c_int64, c_i64p,
        ]
        L.oracle_intersect_all.restype = None
        L.oracle_intersect_all.argtypes = [c_u64p, ctypes.c_int64, c_i64p, ctypes.c_int64, c_i64p]
        _lib = L
    return _lib


# --------------------------------------------------------------------------------------
# FASTA reading, restating pyani_plus/utils.py:40-90 (fasta_bytes_iterator) and
# :142-196 (file_md5sum = md5 of the decompressed bytes).
# --------------------------------------------------------------------------------------
def read_bytes_maybe_gz(path: Path | str) -> bytes:
    raw = Path(path).read_bytes()
    if raw[:2] == b"\x1f\x8b":
        return gzip.decompress(raw)
    return raw


def fasta_records(data: bytes) -> list[tuple[bytes, bytes]]:
    """(title, sequence) per record; whitespace stripped from sequences."""
    records: list[tuple[bytes, bytes]] = []
    title = None
    chunks: list[bytes] = []
    for line in data.splitlines(keepends=True):
        if line[:1] == b">":
            if title is not None:
                records.append((title, b"".join(chunks).translate(None, b" \t\r\n")))
            title = line[1:].rstrip()
            chunks = []
        elif title is not None:
            chunks.append(line.rstrip())
    if title is not None:
        records.append((title, b"".join(chunks).translate(None, b" \t\r\n")))
    return records


def file_md5(path: Path | str) -> str:
    return hashlib.md5(read_bytes_maybe_gz(path)).hexdigest()  # noqa: S324


# --------------------------------------------------------------------------------------
# C-oracle front ends
# --------------------------------------------------------------------------------------
def max_hash(scaled: int) -> int:
    return int(lib().oracle_max_hash(scaled))


def murmur64(key: bytes, seed: int = 42) -> int:
    return int(lib().oracle_murmur64(key, len(key), seed))


def sketch_records(seqs: list[bytes], k: int = 31, scaled: int = 1000, seed: int = 42, *,
                   fast: bool = False) -> np.ndarray:
    """Sorted unique FracMinHash hashes (uint64) of a genome given as its record sequences.

    ``fast=True`` uses the engineered CPU form (the one the baseline timings use) instead of the
    naive checker; both must agree.
    """
    blob = b"".join(seqs)
    offs = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum([len(s) for s in seqs], out=offs[1:])
    cap = max(16, len(blob))  # cannot exceed the number of k-mers
    out = np.empty(cap, dtype=np.uint64)
    buf = ctypes.create_string_buffer(blob, len(blob)) if blob else ctypes.create_string_buffer(1)
    fn = lib().oracle_sketch_fast if fast else lib().oracle_sketch
    n = fn(
        ctypes.cast(buf, ctypes.c_void_p), offs.ctypes.data_as(c_i64p), len(seqs), k,
        max_hash(scaled), seed, out.ctypes.data_as(c_u64p), cap,
    )
    if n < 0:
        msg = f"oracle_sketch failed for k={k}"
        raise ValueError(msg)
    return out[:n].copy()


def sketch_fasta(path: Path | str, k: int = 31, scaled: int = 1000) -> np.ndarray:
    return sketch_records([s for _, s in fasta_records(read_bytes_maybe_gz(path))], k, scaled)


def intersect(a: np.ndarray, b: np.ndarray) -> int:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = np.ascontiguousarray(b, dtype=np.uint64)
    return int(lib().oracle_intersect(a.ctypes.data_as(c_u64p), len(a), b.ctypes.data_as(c_u64p), len(b)))


def ani_from_containment(c: float, k: int = 31) -> float:
    return float(lib().oracle_ani_from_containment(c, k))


def pair_row(ov: int, nq: int, ns: int, k: int = 31) -> dict[str, float] | None:
    """The manysearch row for one ordered pair, or None when branchwater prints no row."""
    out = (ctypes.c_double * 7)()
    if not lib().oracle_pair_row(ov, nq, ns, k, out):
        return None
    keys = (
        "containment", "max_containment", "jaccard", "query_containment_ani",
        "match_containment_ani", "average_containment_ani", "max_containment_ani",
    )
    return dict(zip(keys, (float(x) for x in out), strict=True))


def sig_md5sum(hashes: np.ndarray, k: int = 31) -> str:
    """sourmash's per-sketch md5sum: md5 of str(ksize) followed by each hash in decimal."""
    m = hashlib.md5()  # noqa: S324
    m.update(str(k).encode())
    for h in hashes:
        m.update(str(int(h)).encode())
    return m.hexdigest()


def synth_genome(seed: int, g: int, length: int) -> bytes:
    buf = ctypes.create_string_buffer(length)
    lib().oracle_synth_genome(seed, g, length, ctypes.cast(buf, ctypes.c_void_p))
    return buf.raw


def synth_sketch_batch(seed: int, g0: int, n: int, length: int, k: int, scaled: int,
                       cap: int | None = None) -> tuple[np.ndarray, np.ndarray]:
    """Sketch synthetic genomes g0..g0+n-1 with all host threads; returns (hashes[n,cap], counts[n])."""
    if cap is None:
        cap = int(length / scaled * 1.5) + 256
    out = np.zeros((n, cap), dtype=np.uint64)
    counts = np.zeros(n, dtype=np.int64)
    lib().oracle_synth_sketch_batch(seed, g0, n, length, k, max_hash(scaled),
                                    out.ctypes.data_as(c_u64p), cap, counts.ctypes.data_as(c_i64p))
    if (counts > cap).any():
        msg = "oracle sketch capacity exceeded"
        raise ValueError(msg)
    return out, counts


def intersect_all(hashes: np.ndarray, counts: np.ndarray) -> np.ndarray:
    """All-vs-all intersection counts (int64 n x n, diagonal = sizes) with all host threads."""
    n, cap = hashes.shape
    hashes = np.ascontiguousarray(hashes, dtype=np.uint64)
    counts = np.ascontiguousarray(counts, dtype=np.int64)
    ov = np.zeros((n, n), dtype=np.int64)
    lib().oracle_intersect_all(hashes.ctypes.data_as(c_u64p), cap, counts.ctypes.data_as(c_i64p), n,
                               ov.ctypes.data_as(c_i64p))
    return ov


def num_threads() -> int:
    return int(lib().oracle_num_threads())


# --------------------------------------------------------------------------------------
# Independent pure-Python restatement (tiny inputs only) -- SURVEY.md Appendix B.
# --------------------------------------------------------------------------------------
_M = (1 << 64) - 1


def _rotl(x: int, r: int) -> int:
    return ((x << r) | (x >> (64 - r))) & _M


def _fmix(k: int) -> int:
    k ^= k >> 33
    k = (k * 0xFF51AFD7ED558CCD) & _M
    k ^= k >> 33
    k = (k * 0xC4CEB9FE1A85EC53) & _M
    k ^= k >> 33
    return k


def py_murmur64(key: bytes, seed: int = 42) -> int:
    c1, c2 = 0x87C37B91114253D5, 0x4CF5AD432745937F
    h1 = h2 = seed
    n = len(key)
    nblocks = n // 16
    for i in range(nblocks):
        k1 = int.from_bytes(key[16 * i: 16 * i + 8], "little")
        k2 = int.from_bytes(key[16 * i + 8: 16 * i + 16], "little")
        k1 = (k1 * c1) & _M; k1 = _rotl(k1, 31); k1 = (k1 * c2) & _M; h1 ^= k1  # noqa: E702
        h1 = _rotl(h1, 27); h1 = (h1 + h2) & _M; h1 = (h1 * 5 + 0x52DCE729) & _M  # noqa: E702
        k2 = (k2 * c2) & _M; k2 = _rotl(k2, 33); k2 = (k2 * c1) & _M; h2 ^= k2  # noqa: E702
        h2 = _rotl(h2, 31); h2 = (h2 + h1) & _M; h2 = (h2 * 5 + 0x38495AB5) & _M  # noqa: E702
    tail = key[nblocks * 16:]
    if len(tail) > 8:
        k2 = int.from_bytes(tail[8:], "little")
        k2 = (k2 * c2) & _M; k2 = _rotl(k2, 33); k2 = (k2 * c1) & _M; h2 ^= k2  # noqa: E702
    if len(tail) > 0:
        k1 = int.from_bytes(tail[:8], "little")
        k1 = (k1 * c1) & _M; k1 = _rotl(k1, 31); k1 = (k1 * c2) & _M; h1 ^= k1  # noqa: E702
    h1 ^= n; h2 ^= n  # noqa: E702
    h1 = (h1 + h2) & _M; h2 = (h2 + h1) & _M  # noqa: E702
    h1 = _fmix(h1); h2 = _fmix(h2)  # noqa: E702
    return (h1 + h2) & _M


_COMP = bytes.maketrans(b"ACGT", b"TGCA")


def py_max_hash(scaled: int) -> int:
    if scaled == 0:
        return 0
    if scaled == 1:
        return _M
    return int(float(1 << 64) / float(scaled))


def py_sketch_records(seqs: list[bytes], k: int = 31, scaled: int = 1000, seed: int = 42) -> list[int]:
    mh = py_max_hash(scaled)
    keep: set[int] = set()
    for seq in seqs:
        s = seq.upper()
        for i in range(len(s) - k + 1):
            kmer = s[i: i + k]
            if kmer.strip(b"ACGT"):  # cheap pre-check: ends only; full check next
                continue
            if any(c not in b"ACGT" for c in kmer):
                continue
            rc = kmer.translate(_COMP)[::-1]
            h = py_murmur64(min(kmer, rc), seed)
            if h != 0 and h <= mh:
                keep.add(h)
    return sorted(keep)


def py_ani(c: float, k: int = 31) -> float:
    if c == 0.0:
        return 0.0
    if c == 1.0:
        return 1.0
    return 1.0 - (1.0 - c ** (1.0 / k))
