"""CPU emulation of the K1 thread logic (the same headers the CUDA kernel compiles) vs the oracle.

Runs without a GPU: ``csrc/hostemu.cpp`` walks the kernel's (tile, thread) geometry sequentially using
``kmer_hash.cuh`` / ``pack.cuh`` compiled by g++, so the register-level algorithm (2-bit packing,
reverse-complement by bit tricks, PRMT-based ASCII expansion, canonical choice on packed integers,
specialised murmur) is proven bit-exact before GPU time is spent.  The GPU parity tests
(tests/test_gpu_parity.py) then check the real kernels.
"""

from __future__ import annotations

import ctypes
import os
from pathlib import Path

import numpy as np
import pytest

import __graft_entry__ as entry
from oracle import oracle
from pyani_plus_b200 import stream


@pytest.fixture(scope="module")
def emu() -> ctypes.CDLL:
    entry.build()
    lib = ctypes.CDLL(os.environ.get("PANIB_HOSTEMU_LIB", str(entry.PKG / "libpanib_hostemu.so")))  # override: variant geometry
    lib.emu_sketch_tiles.restype = ctypes.c_int64
    lib.emu_sketch_tiles.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                     ctypes.c_int, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_void_p,
                                     ctypes.c_int64]
    lib.emu_pack_ascii.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
    return lib


def emu_sketch(lib: ctypes.CDLL, genomes: list[list[bytes]], k: int, scaled: int) -> list[np.ndarray]:
    lens = [stream.genome_stream_length(g) for g in genomes]
    toff = stream.plan_tiles(lens)
    buf = np.empty(stream.stream_bytes(toff), dtype=np.uint8)
    stream.fill_ascii_stream(buf, toff, genomes)
    packed = np.zeros(buf.size // 16, dtype=np.uint32)
    mask = np.zeros(buf.size // 32, dtype=np.uint32)
    lib.emu_pack_ascii(buf.ctypes.data, buf.size, packed.ctypes.data, mask.ctypes.data)
    out = []
    for g in range(len(genomes)):
        cap = 1 << 20
        hashes = np.zeros(cap, dtype=np.uint64)
        n = lib.emu_sketch_tiles(packed.ctypes.data, mask.ctypes.data, int(toff[g]), int(toff[g + 1]), k, 42,
                                 oracle.max_hash(scaled), hashes.ctypes.data, cap)
        assert 0 <= n <= cap
        out.append(np.unique(hashes[:n]))
    return out


def _records(path: Path) -> list[bytes]:
    return [s for _, s in oracle.fasta_records(oracle.read_bytes_maybe_gz(path))]


@pytest.mark.parametrize(("k", "scaled"), [(31, 50), (21, 20), (32, 10), (15, 5), (16, 3), (7, 1)])
def test_thread_logic_matches_oracle(emu: ctypes.CDLL, golden: Path, k: int, scaled: int) -> None:
    small = _records(golden / "MIBY01000005.fasta")  # holds a run of 28 N
    large = _records(golden / "MIBY01000011.fasta")
    genomes = [
        _records(golden / "viral_example" / "OP073605.fasta"),
        small,
        [r.lower() for r in large],
        small + large,  # two records: k-mers must not span them
        [b"ACGTNNACGTTTGACCA" * 40, b"", b"TTGACCAGTA" * 50],
        [b"ACGT"],
        [],
    ]
    got = emu_sketch(emu, genomes, k, scaled)
    for g, recs in enumerate(genomes):
        assert got[g].tolist() == oracle.sketch_records(recs, k, scaled).tolist(), (k, g)


def test_thread_logic_viral_sig(emu: ctypes.CDLL, golden: Path) -> None:
    """Straight against a reference .sig file (k=31, scaled=300)."""
    import json

    recs = _records(golden / "viral_example" / "MGV-GENOME-0264574.fas")
    (outer,) = json.loads(
        (golden / "viral_example/intermediates/sourmash/689d3fd6881db36b5e08329cf23cecdd.sig").read_text()
    )
    assert emu_sketch(emu, [recs], 31, 300)[0].tolist() == outer["signatures"][0]["mins"]


def test_tile_boundaries(emu: ctypes.CDLL) -> None:
    base = oracle.synth_genome(20261017, 3, 3 * 4096 + 64)
    genomes = [[base[:n]] for n in (4095, 4096, 4097, 8192, 8192 + 30, 8192 + 31, 30, 31, 32)]
    got = emu_sketch(emu, genomes, 31, 10)
    for g, recs in enumerate(genomes):
        assert got[g].tolist() == oracle.sketch_records(recs, 31, 10).tolist(), g


def test_stream_layout() -> None:
    toff = stream.plan_tiles([0, 4095, 4096, 10])
    assert toff.tolist() == [0, 1, 2, 4, 5]
    buf = np.zeros(stream.stream_bytes(toff), dtype=np.uint8)
    stream.fill_ascii_stream(buf, toff, [[], [b"A" * 4095], [b"C" * 4000, b"G" * 95], [b"T" * 10]])
    assert buf[:4096].tobytes() == b"N" * 4096
    assert buf[4096: 2 * 4096].tobytes() == b"A" * 4095 + b"N"
    assert buf[2 * 4096: 2 * 4096 + 4096].tobytes() == b"C" * 4000 + b"N" + b"G" * 95
    assert buf[3 * 4096: 4 * 4096].tobytes() == b"N" * 4096
    assert buf[4 * 4096: 4 * 4096 + 11].tobytes() == b"T" * 10 + b"N"
    assert (buf[5 * 4096:] == ord("N")).all()
    with pytest.raises(ValueError, match="wrong shape"):
        stream.fill_ascii_stream(buf[:-1], toff, [])


@pytest.mark.parametrize(("k", "scaled", "seed"), [(31, 3, 1), (31, 1, 2), (21, 2, 3), (32, 4, 4), (15, 2, 5), (16, 1, 6),
                                                   (7, 1, 7)])
def test_thread_logic_fuzz(emu: ctypes.CDLL, k: int, scaled: int, seed: int) -> None:
    """Seeded random genomes built to sit on every seam of the kernel's geometry: record and N-run boundaries at
    random offsets (inside a thread's 64-base block, across 16-base packed words, across 4096-base tiles),
    lower case, IUPAC letters, runs shorter and longer than k, records shorter than k, palindromic stretches
    (forward == reverse complement for even k), low-complexity repeats (duplicate hashes).  A small ``scaled``
    keeps every second to every hash, so a single wrong canonical choice or window shows up."""
    rng = np.random.default_rng(20261017 + seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    comp = {65: 84, 67: 71, 71: 67, 84: 65}

    def random_record(length: int) -> bytes:
        seq = acgt[rng.integers(0, 4, length)].copy()
        for _ in range(int(rng.integers(0, 4))):  # N runs / IUPAC letters
            at, run = int(rng.integers(0, max(1, length))), int(rng.choice([1, 2, k - 1, k, k + 1, 40, 70]))
            seq[at: at + run] = rng.choice(np.frombuffer(b"NRYKMnx-", dtype=np.uint8))
        if length > 4 * k and rng.random() < 0.5:  # a reverse-complement palindrome
            at = int(rng.integers(0, length - 2 * k))
            half = seq[at: at + k].copy()
            if all(int(c) in comp for c in half):
                seq[at + k: at + 2 * k] = np.array([comp[int(c)] for c in half[::-1]], dtype=np.uint8)
        if length > 200 and rng.random() < 0.5:  # a tandem repeat: the same k-mers again and again
            at, unit = int(rng.integers(0, length - 150)), int(rng.choice([1, 2, 3, 5]))
            seq[at: at + 120] = np.resize(seq[at: at + unit], 120)
        if rng.random() < 0.3:
            lo = int(rng.integers(0, max(1, length)))
            seq[lo: lo + 50] = np.frombuffer(seq[lo: lo + 50].tobytes().lower(), dtype=np.uint8)
        return seq.tobytes()

    genomes = []
    for _ in range(10):
        n_rec = int(rng.choice([1, 1, 2, 3, 6]))
        recs = []
        for _ in range(n_rec):
            length = int(rng.choice([0, 1, k - 1, k, k + 1, 63, 64, 65, 4095, 4096, 4097,
                                     int(rng.integers(100, 9000))]))
            recs.append(random_record(length))
        genomes.append(recs)
    got = emu_sketch(emu, genomes, k, scaled)
    for g, recs in enumerate(genomes):
        assert got[g].tolist() == oracle.sketch_records(recs, k, scaled).tolist(), (k, seed, g, [len(r) for r in recs])
