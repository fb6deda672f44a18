"""The drop-in CLI path end to end on the GPU: FASTA directory -> .sig cache -> JSON -> SQLite -> matrices.

Modelled on the reference's tests/snakemake/test_sourmash_workflow.py, tests/test_public_cli.py
(:974-1066, :1508-1577), tests/test_coverage.py (:162-174) and tests/test_self_vs_self.py, using the
reference's own fixtures as goldens.
"""

from __future__ import annotations

import csv
import json
import logging
from pathlib import Path

import numpy as np
import pandas as pd
import pytest

from pyani_plus_b200 import db_orm, private_cli, public_cli, setup_logger, tools
from pyani_plus_b200.methods import sourmash
from pyani_plus_b200.utils import file_md5sum

pytestmark = pytest.mark.gpu

KMERSIZE = 31
SCALED = 300  # default scaled=1000 not suitable for the 3 viruses


def compare_sourmash_sig_files(file1: Path, file2: Path) -> bool:
    """Compare two .sig files, ignoring the path part of the filename entry (as the reference does)."""
    data1, data2 = json.loads(file1.read_text()), json.loads(file2.read_text())
    assert isinstance(data1, list) and isinstance(data2, list) and len(data1) == len(data2)
    for entry1, entry2 in zip(data1, data2, strict=True):
        for key in set(entry1) | set(entry2):
            if key == "filename":
                assert Path(entry1[key]).name == Path(entry2[key]).name
            else:
                assert entry1[key] == entry2[key], f"{key} {entry1[key]!r}!={entry2[key]!r}"
    return True


def compare_db_matrices(database_path: Path, matrices_path: Path, absolute_tolerance: float = 2e-8) -> None:
    """One run, one configuration, N^2 comparisons, matrices equal to the golden TSVs (stem labels)."""
    logger = setup_logger(None)
    with db_orm.connect_to_db(logger, database_path) as session:
        (run,) = session.runs()
        assert session.execute("SELECT COUNT(*) FROM configurations").fetchone()[0] == 1
        n = run.genomes.count()
        assert run.comparisons().count() == n**2
        run.cache_comparisons()
        method = run.configuration.method
        md5_to_stem = {_.genome_hash: Path(_.fasta_filename).stem for _ in run.fasta_hashes}
        # the reference uses Path.stem, so NC_002696.fasta.gz -> NC_002696.fasta; goldens use the bare name
        md5_to_stem = {k: v.split(".f")[0] for k, v in md5_to_stem.items()}
        for attr, fname in (("identities", f"{method}_identity.tsv"), ("cov_query", f"{method}_coverage.tsv")):
            expected = pd.read_csv(matrices_path / fname, sep="\t", header=0, index_col=0).sort_index(axis=0).sort_index(axis=1)
            got = getattr(run, attr).rename(index=md5_to_stem, columns=md5_to_stem).sort_index(axis=0).sort_index(axis=1)
            pd.testing.assert_frame_equal(got, expected.astype(float), obj=fname, atol=absolute_tolerance)


def _log_run(tmp_db: Path, fasta: Path, scaled: int = SCALED) -> None:
    tool = tools.get_sourmash()
    private_cli.log_run(
        fasta=fasta, database=tmp_db, cmdline="pyani-plus sourmash ...", status="Testing",
        name="Testing sourmash prepare-genomes", method="sourmash", program=tool.exe_path.name,
        version=tool.version, kmersize=KMERSIZE, extra=f"scaled={scaled}", create_db=True,
    )


def test_sketch_prepare(input_genomes_tiny: Path, tmp_path: Path) -> None:
    """prepare-genomes writes .sig files equal to the fixtures; resume then fills the matrices."""
    cache = tmp_path / "cache"
    cache.mkdir()
    tmp_db = tmp_path / "sig-prepare.db"
    _log_run(tmp_db, input_genomes_tiny)
    private_cli.prepare_genomes(database=tmp_db, run_id=1, cache=cache)
    for expected in (input_genomes_tiny / "intermediates/sourmash").glob("*.sig"):
        generated = cache / f"sourmash_k={KMERSIZE}_scaled={SCALED}" / expected.name
        assert compare_sourmash_sig_files(expected, generated)
    public_cli.resume(database=tmp_db, cache=cache)
    compare_db_matrices(tmp_db, input_genomes_tiny / "matrices")


def test_compute_column_json_and_manysearch_csv(input_genomes_tiny: Path, tmp_path: Path) -> None:
    """compute-column --subject 0 from the FIXTURE .sig cache: JSON rows and manysearch.csv equal the goldens."""
    tmp_db = tmp_path / "col.db"
    _log_run(tmp_db, input_genomes_tiny)
    cache = tmp_path / "cache"
    sig_dir = cache / f"sourmash_k={KMERSIZE}_scaled={SCALED}"
    sig_dir.mkdir(parents=True)
    for sig in (input_genomes_tiny / "intermediates/sourmash").glob("*.sig"):
        (sig_dir / sig.name).write_bytes(sig.read_bytes())
    out = tmp_path / "sourmash.run_1.column_0.json"
    temp = tmp_path / "temp"
    temp.mkdir()
    assert private_cli.compute_column(database=tmp_db, run_id=1, subject="0", json=out, cache=cache, temp=temp,
                                      log=Path("-")) == 0
    golden_rows = {(r["query_name"], r["match_name"]): r for r in csv.DictReader(
        (input_genomes_tiny / "intermediates/sourmash/manysearch.csv").open())}
    data = json.loads(out.read_text())
    assert len(data["comparisons"]) == 9
    for row in data["comparisons"]:
        gold = golden_rows[(row["query_hash"], row["subject_hash"])]
        assert row["identity"] == float(gold["max_containment_ani"])
        assert row["cov_query"] == float(gold["query_containment_ani"])
    ours = {(r["query_name"], r["match_name"]): r for r in csv.DictReader((temp / "c0" / "manysearch.csv").open())}
    assert ours == golden_rows  # every column of every row, as text
    # a single column (1-based over sorted md5) gives that subject's three rows only
    out1 = tmp_path / "col1.json"
    assert private_cli.compute_column(database=tmp_db, run_id=1, subject="2", json=out1, cache=cache,
                                      log=Path("-")) == 0
    subject = sorted(golden_rows)[0][0] if False else sorted({k[0] for k in golden_rows})[1]
    rows1 = json.loads(out1.read_text())["comparisons"]
    assert {r["subject_hash"] for r in rows1} == {subject} and len(rows1) == 3
    for row in rows1:
        assert row["identity"] == float(golden_rows[(row["query_hash"], subject)]["max_containment_ani"])


def test_compute_tile_stale_csv(caplog: pytest.LogCaptureFixture, tmp_path: Path, input_genomes_tiny: Path) -> None:
    query_csv = tmp_path / "query_sigs.csv"
    query_csv.touch()
    subject_csv = tmp_path / "subject_sigs.csv"
    subject_csv.touch()
    logger = setup_logger(None)
    rows = list(sourmash.compute_sourmash_tile(
        logger, tools.get_sourmash(),
        {"689d3fd6881db36b5e08329cf23cecdd", "5584c7029328dc48d33f95f0a78f7e57"},
        {"689d3fd6881db36b5e08329cf23cecdd", "78975d5144a1cd12e98898d573cf6536"},
        input_genomes_tiny / "intermediates/sourmash", tmp_path))
    assert f"Race condition? Replacing intermediate file '{query_csv}'" in caplog.text
    assert f"Race condition? Replacing intermediate file '{subject_csv}'" in caplog.text
    assert len(rows) == 4
    got = {(q, s): (a, b) for q, s, a, b in rows}
    assert got[("689d3fd6881db36b5e08329cf23cecdd", "689d3fd6881db36b5e08329cf23cecdd")] == (1.0, 1.0)
    assert got[("78975d5144a1cd12e98898d573cf6536", "689d3fd6881db36b5e08329cf23cecdd")] == (0.996207756024834, 0.997900938305757)


def test_sourmash_viral_cli(tmp_path: Path, input_genomes_tiny: Path, caplog: pytest.LogCaptureFixture) -> None:
    """pyani-plus sourmash end to end (scaled=300); re-running finds everything already computed."""
    caplog.set_level(logging.INFO)
    tmp_db = tmp_path / "new dir" / "viral ☺.db"
    tmp_db.parent.mkdir()
    assert public_cli.cli_sourmash(database=tmp_db, fasta=input_genomes_tiny, name="Test Run", scaled=300,
                                   create_db=True, cache=tmp_path) == 0
    compare_db_matrices(tmp_db, input_genomes_tiny / "matrices")
    # full precision values as in the reference's scatter-plot TSV (identity, query_cov)
    want = set()
    for line in (input_genomes_tiny / "plots" / "sourmash_query_cov_scatter.tsv").read_text().splitlines()[1:]:
        a, b, _ = line.split("\t")
        want.add((float(a), float(b)))
    logger = setup_logger(None)
    with db_orm.connect_to_db(logger, tmp_db) as session:
        (run,) = session.runs()
        assert run.status == "Done" and run.name == "Test Run"
        assert {(c.identity, c.cov_query) for c in run.comparisons()} == want
        assert run.configuration.program == "panib200" and run.configuration.extra == "scaled=300"
    caplog.clear()
    assert public_cli.cli_sourmash(database=tmp_db, fasta=input_genomes_tiny, name="Again", scaled=300,
                                   cache=tmp_path) == 0
    assert "Database already has all 3²=9 sourmash comparisons" in caplog.text
    out = tmp_path / "export"
    assert public_cli.export_run(database=tmp_db, outdir=out, run_id=1, label="stem") == 0
    exported = pd.read_csv(out / "sourmash_identity.tsv", sep="\t", index_col=0)
    expected = pd.read_csv(input_genomes_tiny / "matrices" / "sourmash_identity.tsv", sep="\t", index_col=0)
    pd.testing.assert_frame_equal(exported, expected)


def test_sourmash_bacteria_gz(tmp_path: Path, input_bacteria: Path) -> None:
    """Gzipped bacterial genomes (one with two records), default scaled=1000 (reference test_public_cli.py:974-990)."""
    tmp_db = tmp_path / "bacteria.sqlite"
    assert public_cli.cli_sourmash(database=tmp_db, fasta=input_bacteria, name="Test Run", create_db=True,
                                   cache=tmp_path) == 0
    compare_db_matrices(tmp_db, input_bacteria / "matrices")
    for expected in (input_bacteria / "intermediates/sourmash").glob("*.sig"):
        assert compare_sourmash_sig_files(expected, tmp_path / "sourmash_k=31_scaled=1000" / expected.name)


def test_bad_alignments_nulls(tmp_path: Path, input_genomes_bad_alignments: Path) -> None:
    """Two phages without a common hash: off-diagonal comparisons are recorded as NULL."""
    tmp_db = tmp_path / "bad.db"
    assert public_cli.cli_sourmash(database=tmp_db, fasta=input_genomes_bad_alignments, scaled=300, create_db=True,
                                   cache=tmp_path) == 0
    compare_db_matrices(tmp_db, input_genomes_bad_alignments / "matrices")
    logger = setup_logger(None)
    with db_orm.connect_to_db(logger, tmp_db) as session:
        (run,) = session.runs()
        rows = list(run.comparisons())
        assert len(rows) == 4
        assert sorted((c.identity is None, c.cov_query is None) for c in rows) == [(False, False)] * 2 + [(True, True)] * 2


def test_coverage_scaled50(tmp_path: Path, golden: Path) -> None:
    """Reference tests/test_coverage.py:54-80,162-174: the cached JSON matrices, literally."""
    seq_dir = tmp_path / "fasta"
    seq_dir.mkdir()
    (seq_dir / "small.fasta").symlink_to(golden / "MIBY01000005.fasta")
    (seq_dir / "large.fasta").symlink_to(golden / "MIBY01000011.fasta")
    (seq_dir / "both.fasta").write_bytes((golden / "MIBY01000005.fasta").read_bytes()
                                         + (golden / "MIBY01000011.fasta").read_bytes())
    tmp_db = tmp_path / "cov.db"
    public_cli.cli_sourmash(database=tmp_db, fasta=seq_dir, name="Artificial", create_db=True, scaled=50,
                            cache=tmp_path)
    checksums = ('["154173fb8e7415ab45532a738572f957","7b6a6226ce00e52edca15565aa0d270d",'
                 '"a0efc718e680e34d2f5c8f5d2286ca9c"]')
    logger = setup_logger(None)
    with db_orm.connect_to_db(logger, tmp_db) as session:
        (run,) = session.runs()
        assert run.df_identity == ("{" f'"columns":{checksums},"index":{checksums},"data":'
                                   "[[1.0,1.0,null],[1.0,1.0,1.0],[null,1.0,1.0]]}")
        assert run.df_cov_query == ("{" f'"columns":{checksums},"index":{checksums},"data":'
                                    "[[1.0,1.0,null],[0.9622440235,1.0,0.9884105907],[null,1.0,1.0]]}")


def test_self_vs_self_single_genome(tmp_path: Path, input_genomes_tiny: Path) -> None:
    """A one-genome run: identity exactly 1.0 (reference tests/test_self_vs_self.py)."""
    seq_dir = tmp_path / "one"
    seq_dir.mkdir()
    (seq_dir / "OP073605.fasta").symlink_to(input_genomes_tiny / "OP073605.fasta")
    tmp_db = tmp_path / "self.db"
    public_cli.cli_sourmash(database=tmp_db, fasta=seq_dir, create_db=True, scaled=300, cache=tmp_path)
    logger = setup_logger(None)
    with db_orm.connect_to_db(logger, tmp_db) as session:
        (run,) = session.runs()
        (comp,) = list(run.comparisons())
        assert comp.identity == 1.0 and comp.cov_query == 1.0


def test_resume_partial_sourmash(caplog: pytest.LogCaptureFixture, capsys: pytest.CaptureFixture[str],
                                 tmp_path: Path, input_genomes_tiny: Path) -> None:
    """A 2x2 run expanded to 3x3: the 4 pre-seeded rows survive (INSERT OR IGNORE), 5 are computed."""
    caplog.set_level(logging.INFO)
    tmp_db = tmp_path / "resume sourmash.sqlite"
    tool = tools.get_sourmash()
    logger = setup_logger(None)
    with db_orm.connect_to_db(logger, tmp_db) as session:
        config = db_orm.db_configuration(session, "sourmash", tool.exe_path.stem, tool.version, kmersize=31,
                                         extra="scaled=300", create=True)
        fasta_to_hash = {f: file_md5sum(f) for f in sorted(input_genomes_tiny.glob("*.f*"))}
        for filename, md5 in fasta_to_hash.items():
            db_orm.db_genome(logger, session, filename, md5, create=True)
        genomes = list(fasta_to_hash.values())
        for query_hash in genomes[:-1]:
            for subject_hash in genomes[:-1]:
                db_orm.db_comparison(session, config.configuration_id, query_hash, subject_hash,
                                     1.0 if query_hash is subject_hash else 0.99)
        db_orm.add_run(session, config, cmdline="pyani-plus sourmash ...", fasta_directory=input_genomes_tiny,
                       status="Partial", name="Test Resuming A Run", fasta_to_hash=fasta_to_hash)
    public_cli.list_runs(database=tmp_db)
    output = capsys.readouterr().out
    assert " 1 analysis runs in " in output, output
    assert " sourmash │    4 │    0 │    5 │  9=3² │ Partial " in output or "Partial" in output, output
    caplog.clear()
    public_cli.resume(database=tmp_db, cache=tmp_path)
    assert "Resuming run-id 1" in caplog.text
    assert "Database already has 4 of 3²=9 sourmash comparisons, 5 needed" in caplog.text
    with db_orm.connect_to_db(logger, tmp_db) as session:
        (run,) = session.runs()
        assert run.status == "Done" and run.comparisons().count() == 9
        kept = [c.identity for c in run.comparisons()
                if c.query_hash in genomes[:-1] and c.subject_hash in genomes[:-1] and c.query_hash != c.subject_hash]
        assert kept == [0.99, 0.99]


def test_resume_errors(tmp_path: Path, input_genomes_tiny: Path) -> None:
    with pytest.raises(SystemExit, match="does not exist"):
        public_cli.resume(database=tmp_path / "none.db")
    tmp_db = tmp_path / "v.db"
    private_cli.log_run(fasta=input_genomes_tiny, database=tmp_db, cmdline="x", status="Testing", name="x",
                        method="sourmash", program="panib200", version="0.0.0-old", kmersize=31,
                        extra="scaled=300", create_db=True)
    with pytest.raises(SystemExit, match=r"We have panib200 version .* but run-id 1 used panib200 version 0\.0\.0-old instead"):
        public_cli.resume(database=tmp_db, cache=tmp_path)


def test_duplicate_genomes_rejected(tmp_path: Path, input_genomes_tiny: Path) -> None:
    seq_dir = tmp_path / "dups"
    seq_dir.mkdir()
    for name in ("a.fasta", "b.fna"):
        (seq_dir / name).write_bytes((input_genomes_tiny / "OP073605.fasta").read_bytes())
    with pytest.raises(SystemExit, match="Multiple genomes with same MD5 checksum 5584c7029328dc48d33f95f0a78f7e57"):
        public_cli.cli_sourmash(database=tmp_path / "d.db", fasta=seq_dir, create_db=True, cache=tmp_path)


def test_synthetic_100_genome_run_matches_engine(tmp_path: Path) -> None:
    """BASELINE config-2 shaped (shorter genomes): the CLI/DB path equals a direct engine computation."""
    from oracle import oracle
    from pyani_plus_b200 import engine

    n, length, scaled = 40, 150_000, 200
    seq_dir = tmp_path / "synthetic"
    seq_dir.mkdir()
    genomes = []
    for g in range(n):
        seq = oracle.synth_genome(20261017, g, length)
        genomes.append([seq])
        (seq_dir / f"g{g:03d}.fna").write_bytes(b">g%d\n" % g + b"\n".join(seq[i:i + 80] for i in range(0, length, 80)) + b"\n")
    tmp_db = tmp_path / "syn.db"
    assert public_cli.cli_sourmash(database=tmp_db, fasta=seq_dir, create_db=True, scaled=scaled, cache=tmp_path) == 0
    want_h, want_c = oracle.synth_sketch_batch(20261017, 0, n, length, 31, scaled)
    want_ov = oracle.intersect_all(want_h, want_c)
    ident, cov = engine.ani_host(want_ov.astype(np.uint32), want_c.astype(np.int32), want_c.astype(np.int32), 31)
    logger = setup_logger(None)
    with db_orm.connect_to_db(logger, tmp_db) as session:
        (run,) = session.runs()
        md5_to_g = {a.genome_hash: int(a.fasta_filename[1:4]) for a in run.fasta_hashes}
        rows = list(run.comparisons())
        assert len(rows) == n * n
        nulls = 0
        for c in rows:
            i, j = md5_to_g[c.query_hash], md5_to_g[c.subject_hash]
            if np.isnan(ident[i, j]):
                assert c.identity is None and c.cov_query is None
                nulls += 1
            else:
                assert c.identity == ident[i, j] and c.cov_query == cov[i, j]
        assert nulls > 0  # distant synthetic genomes share no hash: the NULL path is exercised


def test_bulk_result_path_equals_json_path(tmp_path: Path, input_genomes_tiny: Path, monkeypatch) -> None:
    """The array-backed recording used for large runs writes the same rows as the JSON hand-over,
    including INSERT OR IGNORE on resume of a partial run."""
    db_json, db_bulk = tmp_path / "json.db", tmp_path / "bulk.db"
    assert public_cli.cli_sourmash(database=db_json, fasta=input_genomes_tiny, scaled=300, create_db=True,
                                   cache=tmp_path) == 0
    monkeypatch.setattr(public_cli, "BULK_THRESHOLD", 0)
    assert public_cli.cli_sourmash(database=db_bulk, fasta=input_genomes_tiny, scaled=300, create_db=True,
                                   cache=tmp_path) == 0
    logger = setup_logger(None)
    rows = {}
    for db in (db_json, db_bulk):
        with db_orm.connect_to_db(logger, db) as session:
            (run,) = session.runs()
            assert run.status == "Done"
            rows[db] = sorted((c.query_hash, c.subject_hash, c.identity, c.cov_query, c.aln_length, c.sim_errors,
                               c.cov_subject, c.uname_machine) for c in run.comparisons())
            mats = (run.df_identity, run.df_cov_query, run.df_hadamard)
            rows[str(db)] = mats
    assert rows[db_json] == rows[db_bulk] and len(rows[db_bulk]) == 9
    assert rows[str(db_json)] == rows[str(db_bulk)]
    compare_db_matrices(db_bulk, input_genomes_tiny / "matrices")
