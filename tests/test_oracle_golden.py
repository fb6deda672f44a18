"""Pin the CPU oracle against every golden vector the reference's tests hold for the sourmash path.

These are the checks that make the oracle trustworthy: all 9 fixture ``.sig``
files (hash lists, max_hash, md5sum), all 27 ``manysearch.csv`` rows (intersect_hashes, containment,
jaccard, max_containment and the four ANI columns, compared as exact float values), the matrices,
and the scaled=50 literals of the reference's ``tests/test_coverage.py:162-174``.
"""

from __future__ import annotations

import csv
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle

SETS = {
    "viral_example": 300,
    "bad_alignments": 300,
    "bacterial_example": 1000,
}


def _fasta_files(d: Path) -> list[Path]:
    return sorted(p for p in d.iterdir() if p.is_file() and ".f" in p.name)


def _load_sig(path: Path) -> dict:
    (outer,) = json.loads(path.read_text())
    (inner,) = outer["signatures"]
    return {"outer": outer, **inner}


@pytest.fixture(scope="module")
def sketches(golden: Path) -> dict[str, dict[str, np.ndarray]]:
    """md5 -> sketch for every fixture FASTA, per set."""
    out: dict[str, dict[str, np.ndarray]] = {}
    for name, scaled in SETS.items():
        out[name] = {}
        for fasta in _fasta_files(golden / name):
            out[name][oracle.file_md5(fasta)] = oracle.sketch_fasta(fasta, 31, scaled)
    return out


def test_expected_json_single_kmers(golden: Path) -> None:
    exp = json.loads((golden / "expected.json").read_text())
    for kmer, h in exp["single_kmers_k31_seed42"].items():
        b = kmer.encode()
        rc = b.translate(bytes.maketrans(b"ACGT", b"TGCA"))[::-1]
        canon = min(b, rc)
        assert oracle.murmur64(canon) == h
        assert oracle.py_murmur64(canon) == h
    for scaled, mh in exp["max_hash"].items():
        assert oracle.max_hash(int(scaled)) == mh
        assert oracle.py_max_hash(int(scaled)) == mh
    assert oracle.max_hash(1) == 2**64 - 1
    assert oracle.max_hash(50) == 368934881474191040
    assert oracle.max_hash(100) == 184467440737095520


def test_murmur_c_vs_python_all_lengths() -> None:
    rng = np.random.default_rng(7)
    for n in list(range(0, 70)) + [127, 255]:
        key = bytes(rng.integers(0, 256, n, dtype=np.uint8))
        assert oracle.murmur64(key, 42) == oracle.py_murmur64(key, 42), n
        assert oracle.murmur64(key, 0) == oracle.py_murmur64(key, 0), n


@pytest.mark.parametrize("name", list(SETS))
def test_sig_files_bit_exact(golden: Path, sketches: dict, name: str) -> None:
    """mins / max_hash / md5sum of every fixture .sig (reference: test_sourmash_workflow.py:43-109)."""
    exp = json.loads((golden / "expected.json").read_text())
    sig_dir = golden / name / "intermediates" / "sourmash"
    sigs = sorted(sig_dir.glob("*.sig"))
    assert sigs
    for sig_path in sigs:
        sig = _load_sig(sig_path)
        md5 = sig_path.stem
        assert sig["outer"]["name"] == md5
        got = sketches[name][md5]
        assert sig["max_hash"] == oracle.max_hash(SETS[name])
        assert sig["ksize"] == 31
        assert sig["seed"] == 42
        assert got.tolist() == sig["mins"], f"{name}/{md5}"
        recs = [s for _, s in oracle.fasta_records(oracle.read_bytes_maybe_gz(
            next(f for f in _fasta_files(golden / name) if oracle.file_md5(f) == md5)))]
        assert oracle.sketch_records(recs, 31, SETS[name], fast=True).tolist() == sig["mins"]
        assert len(got) == exp["sketch_sizes"][md5]
        assert oracle.sig_md5sum(got, 31) == sig["md5sum"]
    # every fixture FASTA has a .sig
    assert {p.stem for p in sigs} == set(sketches[name])


@pytest.mark.parametrize("name", list(SETS))
def test_manysearch_rows_exact(golden: Path, sketches: dict, name: str) -> None:
    """Every manysearch.csv row: counts and floats equal as float values (repr-exact)."""
    sk = sketches[name]
    rows = list(csv.DictReader((golden / name / "intermediates" / "sourmash" / "manysearch.csv").open()))
    seen = set()
    for row in rows:
        q, s = row["query_name"], row["match_name"]
        seen.add((q, s))
        ov = oracle.intersect(sk[q], sk[s])
        assert ov == int(row["intersect_hashes"])
        got = oracle.pair_row(ov, len(sk[q]), len(sk[s]), 31)
        assert got is not None
        for key in (
            "containment", "max_containment", "jaccard", "query_containment_ani",
            "match_containment_ani", "average_containment_ani", "max_containment_ani",
        ):
            assert got[key] == float(row[key]), (q, s, key, got[key], row[key])
            assert repr(got[key]) == row[key] or float(repr(got[key])) == float(row[key])
        assert row["query_md5"] == oracle.sig_md5sum(sk[q])
        assert row["match_md5"] == oracle.sig_md5sum(sk[s])
    # pairs with no row are exactly those with zero overlap
    for q in sk:
        for s in sk:
            if (q, s) not in seen:
                assert oracle.intersect(sk[q], sk[s]) == 0
                assert oracle.pair_row(0, len(sk[q]), len(sk[s])) is None
    assert len(rows) == {"viral_example": 9, "bad_alignments": 2, "bacterial_example": 16}[name]


@pytest.mark.parametrize("name", list(SETS))
def test_matrices(golden: Path, sketches: dict, name: str) -> None:
    """identity = max-containment ANI, tolerance as the reference's compare_db_matrices (2e-8)."""
    sk = sketches[name]
    stem_to_md5 = {}
    for fasta in _fasta_files(golden / name):
        stem = fasta.name.split(".")[0] if not fasta.name.startswith("MGV") else fasta.name.rsplit(".", 1)[0]
        stem_to_md5[stem] = oracle.file_md5(fasta)
    lines = (golden / name / "matrices" / "sourmash_identity.tsv").read_text().rstrip("\n").split("\n")
    cols = lines[0].split("\t")[1:]
    for line in lines[1:]:
        cells = line.split("\t")
        q = stem_to_md5[cells[0]]
        for col, cell in zip(cols, cells[1:], strict=True):
            s = stem_to_md5[col]
            row = oracle.pair_row(oracle.intersect(sk[q], sk[s]), len(sk[q]), len(sk[s]))
            if cell == "":
                assert row is None
            else:
                assert row is not None
                assert abs(row["max_containment_ani"] - float(cell)) < 2e-8


def test_coverage_scaled50(golden: Path) -> None:
    """Reference tests/test_coverage.py:162-174: N-run skipping, nulls, 10-digit cov_query."""
    exp = json.loads((golden / "expected.json").read_text())["test_coverage_scaled50"]
    small = (golden / "MIBY01000005.fasta").read_bytes()
    large = (golden / "MIBY01000011.fasta").read_bytes()
    both = small + large
    import hashlib

    files = {hashlib.md5(b).hexdigest(): b for b in (small, large, both)}  # noqa: S324
    assert sorted(files) == exp["md5_sorted"]
    sk = {
        md5: oracle.sketch_records([s for _, s in oracle.fasta_records(data)], 31, 50)
        for md5, data in files.items()
    }
    for data in files.values():  # the engineered (baseline) form agrees with the naive checker
        recs = [s for _, s in oracle.fasta_records(data)]
        for k in (31, 21, 5):
            assert (oracle.sketch_records(recs, k, 50, fast=True).tolist()
                    == oracle.sketch_records(recs, k, 50).tolist())
    assert len(sk["154173fb8e7415ab45532a738572f957"]) == 148  # 149 if N-windows were hashed
    assert len(sk["a0efc718e680e34d2f5c8f5d2286ca9c"]) == 340
    assert len(sk["7b6a6226ce00e52edca15565aa0d270d"]) == 488
    order = exp["md5_sorted"]
    for i, q in enumerate(order):
        for j, s in enumerate(order):
            row = oracle.pair_row(oracle.intersect(sk[q], sk[s]), len(sk[q]), len(sk[s]))
            want_id = exp["df_identity_data"][i][j]
            want_cov = exp["df_cov_query_data"][i][j]
            if want_id is None:
                assert row is None
            else:
                assert row is not None
                assert round(row["max_containment_ani"], 10) == want_id
                assert round(row["query_containment_ani"], 10) == want_cov


def test_python_restatement_agrees_on_small_genome(golden: Path) -> None:
    """The two independent restatements agree on a whole (small) fixture genome, incl. the N run."""
    for fname, scaled in (("MIBY01000005.fasta", 50),):
        recs = [s for _, s in oracle.fasta_records((golden / fname).read_bytes())]
        assert oracle.py_sketch_records(recs, 31, scaled) == oracle.sketch_records(recs, 31, scaled).tolist()


def test_lowercase_and_record_boundaries() -> None:
    rng = np.random.default_rng(3)
    seq = bytes(rng.choice(list(b"ACGT"), 400).astype(np.uint8))
    a = oracle.sketch_records([seq], 21, 5)
    b = oracle.sketch_records([seq.lower()], 21, 5)
    assert a.tolist() == b.tolist() and len(a) > 10
    # splitting into two records loses exactly the k-mers spanning the cut
    two = oracle.sketch_records([seq[:200], seq[200:]], 21, 1)
    one = oracle.sketch_records([seq], 21, 1)
    assert set(two.tolist()) <= set(one.tolist())
    assert len(one) - len(two) <= 20
    assert oracle.py_sketch_records([seq[:200], seq[200:]], 21, 1) == two.tolist()
    # genome shorter than k -> empty sketch -> no self row
    assert len(oracle.sketch_records([b"ACGT"], 31, 1)) == 0
    assert oracle.pair_row(0, 0, 0) is None


def test_synth_generator_properties() -> None:
    g0 = oracle.synth_genome(20261017, 0, 20000)
    g1 = oracle.synth_genome(20261017, 1, 20000)
    assert set(g0) <= set(b"ACGT")
    assert g0 == oracle.synth_genome(20261017, 0, 20000)  # deterministic
    assert g0[:5000] == oracle.synth_genome(20261017, 0, 5000)  # position-addressable
    ident = sum(x == y for x, y in zip(g0, g1, strict=True)) / len(g0)
    assert 0.6 < ident < 0.99
    thr = oracle.lib().oracle_synth_threshold(20261017, 0)
    assert 0.01 * 2**53 <= thr <= 0.2 * 2**53
