"""Parity of the CUDA path (through the C ABI) with the oracle and the reference's golden fixtures.

Bars (BASELINE.json north_star): hashes and intersection counts bit-exact; ANI within 1e-12 absolute
(the host-finalised ANI used by the drop-in path must in fact be exactly equal).
"""

from __future__ import annotations

import csv
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu

SEED = 20261017
ANI_ATOL = 1e-12


@pytest.fixture(scope="module")
def eng():
    from pyani_plus_b200 import engine

    return engine.Engine(0)


def _records(path: Path) -> list[bytes]:
    return [s for _, s in oracle.fasta_records(oracle.read_bytes_maybe_gz(path))]


def _fasta_files(d: Path) -> list[Path]:
    return sorted(p for p in d.iterdir() if p.is_file() and ".f" in p.name)


@pytest.mark.parametrize(("name", "scaled"), [("viral_example", 300), ("bad_alignments", 300),
                                              ("bacterial_example", 1000)])
def test_fixture_sigs_and_manysearch(eng, golden: Path, name: str, scaled: int) -> None:
    """GPU sketches == fixture .sig mins; GPU counts + host ANI == manysearch.csv values exactly."""
    from pyani_plus_b200 import engine

    files = _fasta_files(golden / name)
    md5s = [oracle.file_md5(f) for f in files]
    table = eng.sketch_genomes([_records(f) for f in files], 31, scaled)
    got = dict(zip(md5s, table.to_host(), strict=True))
    for md5, hashes in got.items():
        (outer,) = json.loads((golden / name / "intermediates" / "sourmash" / f"{md5}.sig").read_text())
        (sig,) = outer["signatures"]
        assert hashes.tolist() == sig["mins"]
        assert sig["max_hash"] == engine.max_hash(scaled)
        assert oracle.sig_md5sum(hashes) == sig["md5sum"]
    ov = eng.intersect(table).cpu().numpy()
    counts = table.counts.cpu().numpy()
    ident, cov = engine.ani_host(ov.astype(np.uint32), counts, counts, 31)
    ident_d, cov_d = (t.cpu().numpy() for t in eng.ani_device(eng.intersect(table), table))
    rows = list(csv.DictReader((golden / name / "intermediates" / "sourmash" / "manysearch.csv").open()))
    seen = set()
    for row in rows:
        i, j = md5s.index(row["query_name"]), md5s.index(row["match_name"])
        seen.add((i, j))
        assert ov[i, j] == int(row["intersect_hashes"])
        assert ident[i, j] == float(row["max_containment_ani"])  # exact
        assert cov[i, j] == float(row["query_containment_ani"])  # exact
        assert abs(ident_d[i, j] - float(row["max_containment_ani"])) <= ANI_ATOL
        assert abs(cov_d[i, j] - float(row["query_containment_ani"])) <= ANI_ATOL
    for i in range(len(md5s)):
        for j in range(len(md5s)):
            if (i, j) not in seen:  # no manysearch row <=> NULL
                assert ov[i, j] == 0
                assert np.isnan(ident[i, j]) and np.isnan(cov[i, j])
                assert np.isnan(ident_d[i, j]) and np.isnan(cov_d[i, j])


def test_coverage_scaled50_with_n_run(eng, golden: Path) -> None:
    """Reference tests/test_coverage.py:162-174 (28-N run, two-record genome, nulls)."""
    from pyani_plus_b200 import engine

    exp = json.loads((golden / "expected.json").read_text())["test_coverage_scaled50"]
    small = _records(golden / "MIBY01000005.fasta")
    large = _records(golden / "MIBY01000011.fasta")
    # sorted-md5 order of the reference test: small, both, large
    table = eng.sketch_genomes([small, small + large, large], 31, 50)
    counts = table.counts.cpu().numpy()
    assert counts.tolist() == [148, 488, 340]
    ov = eng.intersect(table).cpu().numpy()
    ident, cov = engine.ani_host(ov.astype(np.uint32), counts, counts, 31)
    for i in range(3):
        for j in range(3):
            want_id, want_cov = exp["df_identity_data"][i][j], exp["df_cov_query_data"][i][j]
            if want_id is None:
                assert np.isnan(ident[i, j])
            else:
                assert round(float(ident[i, j]), 10) == want_id
                assert round(float(cov[i, j]), 10) == want_cov


@pytest.mark.parametrize("k", [31, 21, 7, 15, 16, 32, 33, 51])
def test_kmer_sizes_vs_oracle(eng, golden: Path, k: int) -> None:
    """Fast kernels (k=21,31) and the generic kernel (every other k) vs the oracle; includes N runs,
    lower case and a multi-record genome."""
    genomes = [
        _records(golden / "MIBY01000005.fasta"),
        [r.lower() for r in _records(golden / "MIBY01000011.fasta")],
        _records(golden / "MIBY01000005.fasta") + _records(golden / "MIBY01000011.fasta"),
        _records(golden / "viral_example" / "OP073605.fasta"),
        [b"ACGT" * 3],  # shorter than most k
        [],
        [b"ACGTNNACGTTTGACCA" * 40, b"", b"TTGACCAGTA" * 50],
    ]
    scaled = 20
    got = eng.sketch_genomes(genomes, k, scaled).to_host()
    for g, recs in enumerate(genomes):
        want = oracle.sketch_records(recs, k, scaled)
        assert got[g].tolist() == want.tolist(), (k, g)


def test_scaled_one_keeps_everything(eng, golden: Path) -> None:
    recs = _records(golden / "MIBY01000011.fasta")
    got = eng.sketch_genomes([recs], 31, 1).to_host()[0]
    want = oracle.sketch_records(recs, 31, 1)
    assert got.tolist() == want.tolist()
    assert len(got) > 17000


def test_exact_tile_multiple_and_boundaries(eng) -> None:
    """Genome lengths around the 4096-base tile size; k-mers must not leak across genomes."""
    base = oracle.synth_genome(SEED, 3, 3 * 4096 + 64)
    genomes = [[base[:n]] for n in (4095, 4096, 4097, 8192, 8192 + 30, 8192 + 31, 30, 31, 32, 12288)]
    got = eng.sketch_genomes(genomes, 31, 10).to_host()
    for g, recs in enumerate(genomes):
        assert got[g].tolist() == oracle.sketch_records(recs, 31, 10).tolist(), g


def test_synthetic_family_counts_and_ani(eng) -> None:
    """BASELINE config-2 shaped (scaled down): sketches, full count matrix and ANI vs the oracle."""
    from pyani_plus_b200 import engine

    n, length, k, scaled = 24, 200_000, 31, 100
    genomes = [[oracle.synth_genome(SEED, g, length)] for g in range(n)]
    table = eng.sketch_genomes(genomes, k, scaled)
    want_hashes, want_counts = oracle.synth_sketch_batch(SEED, 0, n, length, k, scaled)
    got = table.to_host()
    for g in range(n):
        assert got[g].tolist() == want_hashes[g, : want_counts[g]].tolist()
    want_ov = oracle.intersect_all(want_hashes, want_counts)
    ov = eng.intersect(table).cpu().numpy()
    assert (ov.astype(np.int64) == want_ov).all()
    counts = table.counts.cpu().numpy()
    ident, cov = engine.ani_host(ov.astype(np.uint32), counts, counts, k)
    ident_d, cov_d = (t.cpu().numpy() for t in eng.ani_device(eng.intersect(table), table))
    for i in range(n):
        for j in range(n):
            row = oracle.pair_row(int(want_ov[i, j]), int(want_counts[i]), int(want_counts[j]), k)
            if row is None:
                assert np.isnan(ident[i, j]) and np.isnan(ident_d[i, j])
            else:
                assert ident[i, j] == row["max_containment_ani"]
                assert cov[i, j] == row["query_containment_ani"]
                assert abs(ident_d[i, j] - row["max_containment_ani"]) <= ANI_ATOL
                assert abs(cov_d[i, j] - row["query_containment_ani"]) <= ANI_ATOL


def test_device_generator_matches_oracle(eng) -> None:
    from pyani_plus_b200 import stream

    n, length = 5, 10_000
    d_ascii, tile_off = eng.synth_ascii_stream(SEED, 7, n, length)
    host = d_ascii.cpu().numpy()
    for g in range(n):
        at = int(tile_off[g]) * stream.TILE
        assert host[at: at + length].tobytes() == oracle.synth_genome(SEED, 7 + g, length)
        assert (host[at + length: int(tile_off[g + 1]) * stream.TILE] == ord("N")).all()


@pytest.mark.parametrize(("cells", "seg_cap"), [(0, 0), (1, 6144), (3, 6144), (7, 2048), (2, 12288)])
def test_intersect_segmentation_variants(eng, cells: int, seg_cap: int) -> None:
    """The multi-cell (large-sketch) path gives the same counts as the single-segment path."""
    rng = np.random.default_rng(11)
    from pyani_plus_b200 import engine

    mh = engine.max_hash(100)
    pool = np.unique(rng.integers(1, mh, 9000, dtype=np.uint64))
    sketches = []
    for i in range(13):
        m = rng.random(len(pool)) < (0.15 + 0.05 * i)
        sketches.append(pool[m])
    sketches.append(np.empty(0, dtype=np.uint64))
    sketches.append(pool[:1])
    table = eng.table_from_host(sketches, 31, 100)
    ov = eng.intersect(table, n_cells=cells, seg_cap=seg_cap).cpu().numpy()
    for i, a in enumerate(sketches):
        for j, b in enumerate(sketches):
            assert ov[i, j] == len(np.intersect1d(a, b)), (i, j)


def test_intersect_rectangular_and_sharded(eng) -> None:
    """Queries x subjects (compute-column shape) and the rank/world split used on several GPUs."""
    rng = np.random.default_rng(5)
    from pyani_plus_b200 import engine

    mh = engine.max_hash(1000)
    pool = np.unique(rng.integers(1, mh, 4000, dtype=np.uint64))
    sk = [pool[rng.random(len(pool)) < 0.5] for _ in range(21)]
    q = eng.table_from_host(sk[:9], 31, 1000)
    s = eng.table_from_host(sk[5:], 31, 1000)
    ov = eng.intersect(q, s).cpu().numpy()
    for i in range(9):
        for j in range(16):
            assert ov[i, j] == len(np.intersect1d(sk[i], sk[5 + j]))
    full = eng.table_from_host(sk, 31, 1000)
    whole = eng.intersect(full).cpu().numpy()
    parts = sum(eng.intersect(full, rank=r, world=3).cpu().numpy().astype(np.int64) for r in range(3))
    assert (parts == whole).all()
    assert (whole == whole.T).all()
    assert (np.diag(whole) == [len(x) for x in sk]).all()


def test_full_size_genome_properties(eng) -> None:
    """BASELINE-size genome (5 Mb, k=31, scaled=1000): size-independent properties.

    * sketch(reverse complement) == sketch(genome)   (canonical k-mers)
    * sketch is sorted, duplicate free, all values in (0, max_hash]
    * a genome vs itself: |A n A| = |A|, ANI exactly 1.0
    * sketch(A) u sketch(B) for a split with a k-1 overlap == sketch(A+B) as one record
    """
    from pyani_plus_b200 import engine

    length = 5_000_000
    seq = oracle.synth_genome(SEED, 0, length)
    rc = seq.translate(bytes.maketrans(b"ACGT", b"TGCA"))[::-1]
    cut = 2_345_678
    table = eng.sketch_genomes([[seq], [rc], [seq[: cut + 30]], [seq[cut:]]], 31, 1000)
    a, b, left, right = table.to_host()
    assert a.tolist() == b.tolist()
    assert (np.diff(a.astype(np.int64)) > 0).all()
    assert a.min() > 0 and a.max() <= engine.max_hash(1000)
    assert abs(len(a) - length / 1000) < 5 * (length / 1000) ** 0.5
    assert np.union1d(left, right).tolist() == a.tolist()
    ov = eng.intersect(table).cpu().numpy()
    assert ov[0, 0] == len(a) and ov[0, 1] == len(a)
    counts = table.counts.cpu().numpy()
    ident, _ = engine.ani_host(ov.astype(np.uint32), counts, counts, 31)
    assert ident[0, 1] == 1.0
    assert ov[2, 3] == len(np.intersect1d(left, right))


def test_config2_full_size_against_oracle(eng) -> None:
    """BASELINE configs[1] at FULL size: 100 synthetic 5 Mb genomes, k=31, scaled=1000.

    Every sketch and the complete 100 x 100 intersection matrix against the oracle (the CPU port runs
    the same workload in a couple of seconds on the host cores), plus ANI within 1e-12 / exactly.
    """
    from pyani_plus_b200 import engine

    n, length, k, scaled = 100, 5_000_000, 31, 1000
    d_ascii, tile_off = eng.synth_ascii_stream(SEED, 0, n, length)
    table = eng.sketch_ascii_stream(d_ascii, tile_off, k, scaled, from_host=False)
    del d_ascii
    want_hashes, want_counts = oracle.synth_sketch_batch(SEED, 0, n, length, k, scaled)
    got = table.to_host()
    for g in range(n):
        assert got[g].tolist() == want_hashes[g, : want_counts[g]].tolist(), g
    want_ov = oracle.intersect_all(want_hashes, want_counts)
    ov = eng.intersect(table).cpu().numpy()
    assert (ov.astype(np.int64) == want_ov).all()
    counts = table.counts.cpu().numpy()
    ident, cov = engine.ani_host(ov.astype(np.uint32), counts, counts, k)
    ident_d, cov_d = (t.cpu().numpy() for t in eng.ani_device(eng.intersect(table), table))
    np.testing.assert_allclose(ident_d, ident, rtol=0, atol=ANI_ATOL, equal_nan=True)
    np.testing.assert_allclose(cov_d, cov, rtol=0, atol=ANI_ATOL, equal_nan=True)
    assert np.isnan(ident).sum() > 0  # distant pairs share no hash (NULL path) ...
    assert (ident[~np.isnan(ident)] > 0.7).all() and (np.diag(ident) == 1.0).all()
    for i, j in ((0, 1), (3, 77), (42, 42), (99, 0)):
        row = oracle.pair_row(int(want_ov[i, j]), int(want_counts[i]), int(want_counts[j]), k)
        if row is None:
            assert np.isnan(ident[i, j])
        else:
            assert ident[i, j] == row["max_containment_ani"] and cov[i, j] == row["query_containment_ani"]


def test_edge_shapes_vs_oracle(eng) -> None:
    """Odd workloads: thousands of tiny genomes, one very large genome (multi-cell K2, >100 buckets),
    repeats (set semantics at insert), all-N and empty genomes, very unequal sketch sizes."""
    from pyani_plus_b200 import engine

    rng = np.random.default_rng(99)
    k, scaled = 31, 100
    big = oracle.synth_genome(SEED, 5, 12_000_000)  # ~120,000 hashes: several K2 cells
    genomes: list[list[bytes]] = [
        [big],
        [big[:6_000_000], big[6_000_000:]],  # same bases as two records (loses 30 spanning k-mers)
        [b"ACGT" * 50_000],  # 4 distinct k-mers, 200,000 windows
        [b"A" * 100_000],
        [b"N" * 5_000],
        [b"ACGTTGCA" * 3 + b"N" + b"GATTACA" * 9],
        [],
    ]
    tiny = [bytes(rng.choice(list(b"ACGT"), 2_000).astype(np.uint8)) for _ in range(1500)]
    genomes += [[t] for t in tiny]
    table = eng.sketch_genomes(genomes, k, scaled)
    got = table.to_host()
    for g in (0, 1, 2, 3, 4, 5, 6, 7, 8, 1506):
        assert got[g].tolist() == oracle.sketch_records(genomes[g], k, scaled, fast=True).tolist(), g
    assert len(got[4]) == 0 and len(got[6]) == 0 and len(got[0]) > 110_000
    assert abs(len(got[0]) - len(got[1])) <= 2
    ov = eng.intersect(table).cpu().numpy()
    n = len(genomes)
    assert ov.shape == (n, n) and (ov == ov.T).all()
    assert (np.diag(ov) == [len(x) for x in got]).all()
    sample = [(0, 1), (0, 2), (1, 0), (0, 7), (7, 8), (2, 3), (4, 4), (0, 1506), (5, 0), (100, 1200)]
    for i, j in sample:
        assert ov[i, j] == oracle.intersect(got[i], got[j]), (i, j)
    counts = table.counts.cpu().numpy()
    ident, cov = engine.ani_host(ov.astype(np.uint32), counts, counts, k)
    assert np.isnan(ident[4, 4]) and np.isnan(ident[6, 6])  # empty sketches: no self row
    assert ident[0, 0] == 1.0 and ident[0, 1] == ident[1, 0] and cov[1, 0] >= cov[0, 1]


def test_step_pipeline_eager_and_graph(eng) -> None:
    """pipeline.SourmashStep (K1 -> finalize -> K2 -> ANI with no read-back in between): the eager step,
    the replayed CUDA graph and the host-input (H2D inside) form give the oracle's sketches and counts."""
    import torch

    from pyani_plus_b200 import pipeline

    n, length, k, scaled = 12, 400_000, 31, 200
    d_ascii, tile_off = eng.synth_ascii_stream(SEED, 40, n, length)
    plan = eng.plan_stream(tile_off, scaled)
    bufs = eng.alloc_stream_buffers(plan, ascii_too=True, host_packed=True)
    tab = eng.alloc_table(plan)
    eng.pack(d_ascii, plan, bufs)
    h_ascii = torch.empty(plan.n_bases, dtype=torch.uint8)
    h_ascii.copy_(d_ascii)
    want_h, want_c = oracle.synth_sketch_batch(SEED, 40, n, length, k, scaled)
    want_ov = oracle.intersect_all(want_h, want_c)
    stepper = pipeline.SourmashStep(eng, plan, bufs, tab, k, h_ascii=h_ascii)

    def check(out: dict) -> None:
        got = out["table"].to_host()
        for g in range(n):
            assert got[g].tolist() == want_h[g, : want_c[g]].tolist(), g
        assert (out["ov"].cpu().numpy().astype(np.int64) == want_ov).all()

    check(stepper.run())
    out = stepper.run(from_host=True, to_host=True)
    check(out)
    ident_eager = out["identity_host"].clone()
    assert not stepper.capture(from_host=True, to_host=True)  # host work inside: not a graph
    assert stepper.capture(to_host=True)
    tab["table"].fill_(7)  # stale rows must be overwritten by the replay
    for _ in range(2):
        out = stepper.replay(to_host=True)
        stepper.finish()
    check(out)
    np.testing.assert_array_equal(out["identity_host"].numpy(), ident_eager.numpy())
    # the older host form (ASCII over PCIe, packed on the device) still gives the same rows
    eng.sketch_ascii_host(h_ascii.pin_memory(), plan, bufs, tab, k)
    assert eng.check_status() == 0
    from pyani_plus_b200 import engine as _eng
    got = _eng.SketchTable(tab["table"], tab["counts"], k, scaled).to_host()
    for g in range(n):
        assert got[g].tolist() == want_h[g, : want_c[g]].tolist(), g
    # already-packed host buffers (h_ascii = None): copy + K1 only (after a host-input step h_packed / h_mask
    # are scratch, so fill them with the dense packed form first)
    hp, hm = _eng.pack_host(h_ascii.numpy())
    bufs["h_packed"].copy_(torch.from_numpy(hp.view(np.int32)))
    bufs["h_mask"].copy_(torch.from_numpy(hm.view(np.int32)))
    tab["table"].fill_(7)
    eng.sketch_host(None, plan, bufs, tab, k)
    assert eng.check_status() == 0
    got = _eng.SketchTable(tab["table"], tab["counts"], k, scaled).to_host()
    for g in range(n):
        assert got[g].tolist() == want_h[g, : want_c[g]].tolist(), g
    ident, _ = engine_ani_host(out["ov"], out["table"], k)
    np.testing.assert_allclose(out["identity_host"].numpy(), ident, rtol=0, atol=ANI_ATOL, equal_nan=True)
    # the inverted-index K2 inside the step, eager and captured (CUB sort + scan inside the graph)
    idx = pipeline.SourmashStep(eng, plan, bufs, tab, k, k2_method="index")
    check(idx.run())
    assert eng.last_intersect_method == "index"
    assert idx.capture()
    for _ in range(2):
        out = idx.replay()
        idx.finish()
    check(out)
    auto = pipeline.SourmashStep(eng, plan, bufs, tab, k)  # "auto" is resolved on the first step
    check(auto.run())
    assert auto.k2_method in ("probe", "index")
    # a size hint that is too small is reported, not silently truncated
    small = pipeline.SourmashStep(eng, plan, bufs, tab, k, size_hint=100)
    with pytest.raises(Exception, match="size_hint"):
        small.run()


def engine_ani_host(ov, table, k: int):
    from pyani_plus_b200 import engine

    counts = table.counts.cpu().numpy()
    return engine.ani_host(ov.cpu().numpy().astype(np.uint32), counts, counts, k)


def _pool_sketches(rng, n: int, pool: int, lo: int, hi: int) -> list[np.ndarray]:
    """n sketches drawn from a shared pool of hashes (so that hashes are shared by 1..n genomes)."""
    universe = np.unique(rng.integers(1, 1 << 53, pool, dtype=np.uint64))
    weights = rng.random(len(universe)) ** 3  # a few very frequent hashes, many rare ones
    weights /= weights.sum()
    out = []
    for _ in range(n):
        size = int(rng.integers(lo, hi + 1))
        out.append(np.sort(rng.choice(universe, size=min(size, len(universe)), replace=False, p=weights)))
    return out


@pytest.mark.parametrize("tau", [2, 5, 0, 100000])
def test_intersect_index_form_equals_probe_and_oracle(eng, tau: int) -> None:
    """The inverted-index form of K2 (sort + bit matrix + rare-pair adds) gives the probing kernel's and
    the oracle's counts: all runs frequent (tau=2), mixed, the default, all runs rare (huge tau); with
    empty sketches, identical sketches, single-hash sketches; and sharded over 3 ranks."""
    rng = np.random.default_rng(7 + tau)
    sk = _pool_sketches(rng, 70, 6000, 0, 900)
    sk[3] = np.zeros(0, dtype=np.uint64)
    sk[10] = sk[11].copy()  # identical genomes
    sk[20] = sk[21][:1].copy()
    sk.append(np.array([5, 9, (1 << 53) + 12345], dtype=np.uint64))
    table = eng.table_from_host(sk, 31, 1000)
    n = len(sk)
    want = np.zeros((n, n), dtype=np.int64)
    for i in range(n):
        for j in range(n):
            want[i, j] = len(np.intersect1d(sk[i], sk[j], assume_unique=True))
    probe = eng.intersect(table, method="probe").cpu().numpy().astype(np.int64)
    assert (probe == want).all()
    got = eng.intersect(table, method="index", tau=tau).cpu().numpy().astype(np.int64)
    assert eng.last_intersect_method == "index"
    assert (got == want).all(), np.argwhere(got != want)[:5]
    # too small a capacity hint is detected and redone with the real maximum
    again = eng.intersect(table, method="index", tau=tau, max_count=50).cpu().numpy().astype(np.int64)
    assert (again == want).all()
    parts = sum(eng.intersect(table, method="index", tau=tau, rank=r, world=3).cpu().numpy().astype(np.int64)
                for r in range(3))
    assert (parts == want).all()


def test_intersect_auto_picks_a_method_and_matches(eng) -> None:
    """method="auto" on a family large enough for the index to be considered: same counts either way."""
    n, length, k, scaled = 150, 600_000, 31, 100
    d_ascii, tile_off = eng.synth_ascii_stream(SEED, 300, n, length)
    table = eng.sketch_ascii_stream(d_ascii, tile_off, k, scaled, from_host=False)
    del d_ascii
    probe = eng.intersect(table, method="probe").cpu().numpy()
    auto = eng.intersect(table).cpu().numpy()
    assert eng.last_intersect_method in ("probe", "index")
    index = eng.intersect(table, method="index").cpu().numpy()
    assert (auto == probe).all() and (index == probe).all()
    assert (np.diag(index) == table.counts.cpu().numpy()).all()
    # rectangular calls cannot use the index
    with pytest.raises(ValueError, match="all-vs-all"):
        eng.intersect(table, table, method="index")


def test_survivor_workspace_and_direct_insert_agree(eng, monkeypatch: pytest.MonkeyPatch) -> None:
    """K1 parks survivors in the registered workspace and a second kernel inserts them; without a workspace
    (or when a CTA's region overflows) it inserts directly from the hashing loop.  Same sketches either way,
    also for dense sketches (scaled=20: every CTA region overflows into the direct path) and equal to the oracle."""
    from pyani_plus_b200 import engine as _eng

    n, length, k = 5, 600_000, 31
    d_ascii, tile_off = eng.synth_ascii_stream(SEED, 7, n, length)
    for scaled in (1000, 20):
        with_ws = eng.sketch_ascii_stream(d_ascii, tile_off, k, scaled, from_host=False).to_host()
        assert eng.device.index in _eng._WORKSPACES  # noqa: SLF001
        _eng._check(eng.lib.panib_set_workspace(None, 0))  # noqa: SLF001
        monkeypatch.setattr(eng, "_ensure_workspace", lambda plan: None)
        direct = eng.sketch_ascii_stream(d_ascii, tile_off, k, scaled, from_host=False).to_host()
        monkeypatch.undo()
        _eng._WORKSPACES.pop(eng.device.index, None)  # noqa: SLF001  the next alloc_table registers a fresh one
        # a workspace far too small for the survivors: regions overflow, the rest is inserted directly
        import torch

        tiny = torch.empty(1 << 16, dtype=torch.uint8, device=eng.device)
        _eng._check(eng.lib.panib_set_workspace(tiny.data_ptr(), tiny.numel()))  # noqa: SLF001
        monkeypatch.setattr(eng, "_ensure_workspace", lambda plan: None)
        overflow = eng.sketch_ascii_stream(d_ascii, tile_off, k, scaled, from_host=False).to_host()
        monkeypatch.undo()
        _eng._check(eng.lib.panib_set_workspace(None, 0))  # noqa: SLF001
        want_h, want_c = oracle.synth_sketch_batch(SEED, 7, n, length, k, scaled)
        for g in range(n):
            want = want_h[g, : want_c[g]].tolist()
            assert with_ws[g].tolist() == want and direct[g].tolist() == want and overflow[g].tolist() == want, (scaled, g)


def test_ingest_pipeline_paths_agree(eng, monkeypatch: pytest.MonkeyPatch) -> None:
    """``panib_sketch_packed_host``: whichever way a chunk travels -- packed by the host threads with a sparse
    or a dense validity mask, or as plain ASCII from the tail of the stream and packed by the GPU -- the device
    ends up with the words and mask bits of the device pack kernel and with the oracle's sketches.  Genomes
    with N runs, lower case and several records make many tiles dirty; one genome of only N makes a chunk
    (nearly) all dirty, which takes the dense route inside the sparse form."""
    import torch

    from pyani_plus_b200 import engine as _eng

    rng = np.random.default_rng(5)
    k, scaled = 31, 100
    genomes = []
    for g in range(9):
        recs = []
        for _ in range(1 + g % 3):
            seq = bytearray(oracle.synth_genome(SEED, 300 + g, 700_000 + 4096 * g))
            for _ in range(g * 40):  # scattered N runs and lower-case stretches
                at = int(rng.integers(0, len(seq) - 200))
                n_run = int(rng.integers(1, 90))
                seq[at: at + n_run] = b"N" * n_run
                lo = int(rng.integers(0, len(seq) - 200))
                seq[lo: lo + 50] = bytes(seq[lo: lo + 50]).lower()
            recs.append(bytes(seq))
        genomes.append(recs)
    genomes.insert(4, [b"N" * 5_000_000])  # covers a whole chunk with nothing but invalid bases
    from pyani_plus_b200 import stream as pstream

    tile_off = pstream.plan_tiles([pstream.genome_stream_length(recs) for recs in genomes])
    plan = eng.plan_stream(tile_off, scaled)
    h_pageable = torch.empty(plan.n_bases, dtype=torch.uint8)
    pstream.fill_ascii_stream(h_pageable.numpy(), tile_off, genomes)
    h_pinned = h_pageable.pin_memory()
    bufs = eng.alloc_stream_buffers(plan, ascii_too=True, host_packed=True)
    tab = eng.alloc_table(plan)
    eng.pack(h_pinned.to(eng.device), plan, bufs)
    want_packed, want_mask = bufs["packed"].clone(), bufs["mask"].clone()
    want = [oracle.sketch_records(recs, k, scaled) for recs in genomes]

    def run(h_ascii, threads: int = 0, **env: str) -> None:
        for key, val in env.items():
            monkeypatch.setenv(key, val)
        bufs["packed"].fill_(-1)
        bufs["mask"].fill_(0x55555555)
        tab["table"].fill_(7)
        eng.sketch_host(h_ascii, plan, bufs, tab, k, threads=threads)
        assert eng.check_status() == 0
        for key in env:
            monkeypatch.delenv(key)
        assert torch.equal(bufs["packed"], want_packed), env
        assert torch.equal(bufs["mask"], want_mask), env
        got = _eng.SketchTable(tab["table"], tab["counts"], k, scaled).to_host()
        for g, w in enumerate(want):
            assert got[g].tolist() == w.tolist(), (env, g)

    run(h_pageable)                                  # sparse mask, host threads only (pageable ASCII)
    run(h_pinned)                                    # + raw chunks whenever the link would idle
    run(h_pinned, PANIB_INGEST_RAW="2")              # raw chunks from the tail as long as any is free
    run(h_pinned, PANIB_INGEST_RAW="0")              # sparse mask only
    run(h_pinned, PANIB_INGEST_SPARSE="0")           # dense mask, in order
    # packed words through a small ring inside h_packed (1 MB = 16 blocks here, so that it wraps several times)
    run(h_pageable, PANIB_INGEST_RING_MB="1")
    run(h_pinned, PANIB_INGEST_RING_MB="1")
    run(h_pinned, PANIB_INGEST_RING_MB="1", PANIB_INGEST_RAW="2")
    run(h_pinned, PANIB_INGEST_RING_MB="1", PANIB_INGEST_SPARSE="0")
    run(h_pinned, PANIB_INGEST_RING_MB="0")          # the whole-stream buffer with streaming stores
    run(h_pinned, 1, PANIB_INGEST_RING_MB="1")       # one host thread: it packs itself, so no ring (it would wait for itself)
    run(h_pinned, 2, PANIB_INGEST_RING_MB="1")       # the submitting thread and one worker
    scratch = bufs.pop("ingest_scratch")             # no device scratch at all: dense mask, in order
    run(h_pinned)
    run(h_pinned, PANIB_INGEST_RING_MB="1")
    bufs["ingest_scratch"] = scratch[: scratch.numel() // 8]  # room for the mask ring only
    run(h_pinned, PANIB_INGEST_RAW="2")


@pytest.mark.parametrize("k", [21, 31])
def test_other_seeds_take_the_general_instantiation(eng, k: int) -> None:
    """The fast K1 kernels exist twice: seed 42 (sourmash's; constants as immediates) and any other seed
    (constants in uniform registers).  Both equal the oracle."""
    genomes = [[oracle.synth_genome(SEED, 900 + g, 150_000 + 4096 * g)] for g in range(3)]
    for seed in (42, 7, 0xFFFFFFFF):
        got = eng.sketch_genomes(genomes, k, 50, seed=seed).to_host()
        for g, recs in enumerate(genomes):
            assert got[g].tolist() == oracle.sketch_records(recs, k, 50, seed).tolist(), (seed, g)
