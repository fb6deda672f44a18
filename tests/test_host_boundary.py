"""Host side of the drop-in boundary, CPU only: ORM records, JSON hand-over, .sig files, parser, errors.

Modelled on the reference's tests/test_sourmash.py, tests/test_json.py, tests/test_orm.py and
tests/test_utils.py for the sourmash path (same scenarios, same expected messages).
"""

from __future__ import annotations

import gzip
import json
import sqlite3
from pathlib import Path

import numpy as np
import pytest

from pyani_plus_b200 import db_orm, private_cli, setup_logger, sigfile, tools, utils
from pyani_plus_b200.methods import sourmash


# ------------------------------------------------------------------------------------- utils
def test_fasta_iterator_and_md5(golden: Path, tmp_path: Path) -> None:
    fasta = golden / "viral_example" / "OP073605.fasta"
    with fasta.open("rb") as handle:
        ((title, seq),) = list(utils.fasta_bytes_iterator(handle))
    assert title == b"OP073605.1 MAG: Bacteriophage sp. isolate 0984_12761, complete genome"
    assert len(seq) == 57793
    assert utils.file_md5sum(fasta) == "5584c7029328dc48d33f95f0a78f7e57"
    # gzip: md5 of the decompressed contents, records parsed transparently
    assert utils.file_md5sum(golden / "bacterial_example" / "NC_011916.fas.gz") == "9d72a8fb513cf9cc8cc6605a0ad4e837"
    recs = utils.read_fasta_records(golden / "bacterial_example" / "NC_002696.fasta.gz")
    assert len(recs) == 2 and sum(len(s) for _, s in recs) == 4016947
    with fasta.open() as handle, pytest.raises(ValueError, match="requires a handle in binary mode"):
        next(utils.fasta_bytes_iterator(handle))  # type: ignore[arg-type]
    with pytest.raises(ValueError, match="not found"):
        utils.file_md5sum(tmp_path / "missing.fasta")
    # mangled input: leading junk, CRs, blank lines, internal spaces
    messy = tmp_path / "messy.fasta"
    messy.write_bytes(b"junk\n\n>one desc  \r\nAC GT\r\n\r\nNN\n>two\nacgt\n")
    with messy.open("rb") as handle:
        assert list(utils.fasta_bytes_iterator(handle)) == [(b"one desc", b"ACGTNN"), (b"two", b"acgt")]
    assert utils.filename_stem("relative/path/example.fna.gz") == "example"


def test_check_fasta_and_db(tmp_path: Path, golden: Path) -> None:
    logger = setup_logger(None)
    assert len(utils.check_fasta(logger, golden / "viral_example")) == 3
    assert len(utils.check_fasta(logger, golden / "bacterial_example")) == 4
    with pytest.raises(SystemExit, match="is not a directory"):
        utils.check_fasta(logger, tmp_path / "nope")
    with pytest.raises(SystemExit, match="No FASTA input genomes under"):
        utils.check_fasta(logger, tmp_path)
    with pytest.raises(SystemExit, match="does not exist, but not using --create-db"):
        utils.check_db(logger, tmp_path / "new.db", create_db=False)
    utils.check_db(logger, tmp_path / "new.db", create_db=True)


# ------------------------------------------------------------------------------------- ORM
def _setup_run(tmp_db: Path, fasta: Path, **config) -> None:
    private_cli.log_run(
        fasta=fasta, database=tmp_db, cmdline="pyani-plus sourmash ...", status="Testing",
        name="Testing sourmash prepare-genomes", create_db=True, **config,
    )


def test_schema_matches_reference(tmp_path: Path) -> None:
    """Table and column names / nullability as SQLAlchemy creates them for pyani_plus/db_orm.py."""
    logger = setup_logger(None)
    db = tmp_path / "schema.db"
    db_orm.connect_to_db(logger, db).close()
    con = sqlite3.connect(db)
    tables = {r[0] for r in con.execute("SELECT name FROM sqlite_master WHERE type='table'")}
    assert tables == {"genomes", "configurations", "runs", "comparisons", "runs_genomes"}
    cols = {t: [(r[1], r[2], r[3], r[5]) for r in con.execute(f"PRAGMA table_info({t})")] for t in tables}
    assert cols["genomes"] == [("genome_hash", "VARCHAR", 1, 1), ("path", "VARCHAR", 1, 0),
                               ("length", "INTEGER", 1, 0), ("description", "VARCHAR", 1, 0)]
    assert [c[0] for c in cols["comparisons"]] == [
        "comparison_id", "query_hash", "subject_hash", "configuration_id", "identity", "aln_length",
        "sim_errors", "cov_query", "cov_subject", "uname_system", "uname_release", "uname_machine"]
    assert [c[0] for c in cols["runs"]] == [
        "run_id", "configuration_id", "cmdline", "fasta_directory", "date", "status", "name", "df_identity",
        "df_cov_query", "df_aln_length", "df_sim_errors", "df_hadamard"]
    sql = " ".join(r[0] for r in con.execute("SELECT sql FROM sqlite_master WHERE type='table'"))
    for name in ("pk_genomes", "uq_configurations_method", "uq_comparisons_query_hash",
                 "fk_comparisons_query_hash_genomes", "fk_runs_configuration_id_configurations",
                 "pk_runs_genomes", "fk_runs_genomes_run_id_runs"):
        assert name in sql


def test_orm_records_and_insert_or_ignore(tmp_path: Path, input_genomes_tiny: Path) -> None:
    logger = setup_logger(None)
    db = tmp_path / "orm.db"
    with db_orm.connect_to_db(logger, db) as session:
        with pytest.raises(db_orm.NoResultFound, match="Requested configuration not already in DB"):
            db_orm.db_configuration(session, "sourmash", "panib200", "0.1.0", kmersize=31, extra="scaled=300")
        config = db_orm.db_configuration(session, "sourmash", "panib200", "0.1.0", kmersize=31, extra="scaled=300",
                                         create=True)
        again = db_orm.db_configuration(session, "sourmash", "panib200", "0.1.0", kmersize=31, extra="scaled=300")
        assert config.configuration_id == again.configuration_id == 1
        assert repr(config) == ("Configuration(configuration_id=1, program='panib200', version='0.1.0', "
                                "fragsize=None, mode=None, kmersize=31, minmatch=None, extra='scaled=300')")
        fasta_to_hash = {f: utils.file_md5sum(f) for f in sorted(input_genomes_tiny.glob("*.f*"))}
        with pytest.raises(db_orm.NoResultFound, match="Requested genome not already in DB"):
            db_orm.db_genome(logger, session, next(iter(fasta_to_hash)), "x" * 32)
        for f, md5 in fasta_to_hash.items():
            g = db_orm.db_genome(logger, session, f, md5, create=True)
        assert g.length == 57793 and g.description.startswith("OP073605.1 MAG")
        run = db_orm.add_run(session, config, "pyani-plus sourmash ...", input_genomes_tiny, "Partial", "A run",
                             fasta_to_hash=fasta_to_hash)
        assert run.run_id == 1 and run.genomes.count() == 3 and run.fasta_hashes.count() == 3
        assert run.comparisons().count() == 0
        hashes = sorted(fasta_to_hash.values())
        first = db_orm.db_comparison(session, 1, hashes[0], hashes[1], identity=0.5, cov_query=0.25)
        assert db_orm.db_comparison(session, 1, hashes[0], hashes[1], identity=0.9).identity == 0.5  # not altered
        assert first.comparison_id == 1
        entries = [{"query_hash": q, "subject_hash": s, "identity": 1.0 if q == s else None, "cov_query": None,
                    "configuration_id": 1, "uname_system": "L", "uname_release": "r", "uname_machine": "m"}
                   for q in hashes for s in hashes]
        assert db_orm.insert_comparisons_with_retries(logger, session, entries)
        assert run.comparisons().count() == 9
        assert run.comparisons().where_subject(hashes[1]).count() == 3
        kept = [c for c in run.comparisons() if (c.query_hash, c.subject_hash) == (hashes[0], hashes[1])]
        assert kept[0].identity == 0.5 and kept[0].cov_query == 0.25  # INSERT OR IGNORE kept the old row
        run.cache_comparisons()
        run.status = "Done"
        session.commit()
    with db_orm.connect_to_db(logger, db) as session:
        run = db_orm.load_run(session, check_complete=True)
        assert run.status == "Done"
        ident = run.identities
        assert list(ident.index) == hashes and ident.loc[hashes[0], hashes[1]] == 0.5
        assert np.isnan(ident.loc[hashes[1], hashes[0]])
        assert json.loads(run.df_identity)["columns"] == hashes
        assert run.hadamard.loc[hashes[0], hashes[1]] == 0.125
        assert list(run.relabelled_matrix(run.identities, "stem").index) == sorted(
            utils.filename_stem(f.name) for f in fasta_to_hash)
        with pytest.raises(ValueError, match="Unexpected label scheme 'unknown'"):
            run.relabelled_matrix(run.identities, "unknown")
        with pytest.raises(SystemExit, match="Database has no run-id 7"):
            db_orm.load_run(session, 7)


def test_db_genome_gzip_name_checks(tmp_path: Path) -> None:
    logger = setup_logger(None)
    plain = tmp_path / "plain.fasta.gz"
    plain.write_bytes(b">x\nACGT\n")
    zipped = tmp_path / "zipped.fasta"
    zipped.write_bytes(gzip.compress(b">x\nACGT\n"))
    with db_orm.connect_to_db(logger, ":memory:") as session:
        with pytest.raises(SystemExit, match=r"Has \.gz ending, but plain\.fasta\.gz is NOT gzip compressed"):
            db_orm.db_genome(logger, session, plain, "a" * 32, create=True)
        with pytest.raises(SystemExit, match=r"No \.gz ending, but zipped\.fasta is gzip compressed"):
            db_orm.db_genome(logger, session, zipped, "b" * 32, create=True)
        with pytest.raises(SystemExit, match="Database contains no runs"):
            db_orm.load_run(session)


# ------------------------------------------------------------------------------------- JSON
def test_json_export_import_round_trip(tmp_path: Path, input_genomes_tiny: Path) -> None:
    logger = setup_logger(None)
    db = tmp_path / "json.db"
    tool = tools.get_sourmash()
    _setup_run(db, input_genomes_tiny, method="sourmash", program=tool.exe_path.stem, version=tool.version,
               kmersize=31, extra="scaled=300")
    out = tmp_path / "column_0.json"
    with db_orm.connect_to_db(logger, db) as session:
        run = db_orm.load_run(session, 1)
        hashes = sorted(a.genome_hash for a in run.fasta_hashes)
        entries = [{"query_hash": q, "subject_hash": s, "identity": 0.99 if q != s else 1.0,
                    "cov_query": None if q < s else 0.5, "configuration_id": 1, "uname_system": "x",
                    "uname_release": "y", "uname_machine": "z"} for q in hashes for s in hashes]
        private_cli.export_json_db_entries(logger, out, run.configuration, entries)
        data = json.loads(out.read_text())
        assert set(data) == {"configuration", "uname", "comparisons"}
        assert data["configuration"] == {"method": "sourmash", "program": tool.exe_path.stem,
                                         "version": tool.version, "fragsize": None, "mode": None, "kmersize": 31,
                                         "minmatch": None, "extra": "scaled=300"}
        assert set(data["comparisons"][0]) == {"query_hash", "subject_hash", "identity", "cov_query"}
        assert private_cli.import_json_comparisons(logger, session, out) == 9
        assert run.comparisons().count() == 9
        assert private_cli.import_json_comparisons(logger, session, out) == 9  # idempotent
        assert run.comparisons().count() == 9


@pytest.mark.parametrize(("text", "message"), [
    ("{}", "does not use the expected structure"),
    ("[1,2", "invalid"),
    ('{"configuration": {}, "uname": {}, "comparisons": []}', "uname incomplete"),
    ('{"configuration": {"method": "x"}, "uname": {"system": 1, "release": 2, "machine": 3}, "comparisons": []}',
     "configuration incomplete"),
    ('{"configuration": {"method": "x", "program": "y", "version": "z"}, "uname": {"system": 1, "release": 2, '
     '"machine": 3}, "comparisons": []}', "configuration not in database"),
])
def test_json_import_errors(tmp_path: Path, text: str, message: str) -> None:
    logger = setup_logger(None)
    bad = tmp_path / "bad.json"
    bad.write_text(text)
    with db_orm.connect_to_db(logger, ":memory:") as session, pytest.raises(SystemExit, match=message):
        private_cli.import_json_comparisons(logger, session, bad)
    bad.write_text("")
    with db_orm.connect_to_db(logger, ":memory:") as session:
        assert private_cli.import_json_comparisons(logger, session, bad) == 0


# ------------------------------------------------------------------------------------- .sig files
@pytest.mark.parametrize("name", ["viral_example", "bad_alignments", "bacterial_example"])
def test_sig_write_reproduces_fixture_files(golden: Path, tmp_path: Path, name: str) -> None:
    """Re-writing a fixture's sketch gives a file equal key for key (reference: test_sourmash_workflow.py:43-67)."""
    for sig_path in sorted((golden / name / "intermediates" / "sourmash").glob("*.sig")):
        want = json.loads(sig_path.read_text())
        sig = sigfile.read_sig(sig_path, ksize=31)
        assert sig["name"] == sig_path.stem and sig["hashes"].dtype == np.uint64
        assert sigfile.sketch_md5sum(sig["hashes"], 31) == sig["md5sum"]
        out = tmp_path / sig_path.name
        sigfile.write_sig(out, filename=want[0]["filename"], name=sig["name"], ksize=31,
                          max_hash=sig["max_hash"], hashes=sig["hashes"])
        got = json.loads(out.read_text())
        assert got == want
        assert list(got[0]) == list(want[0]) and list(got[0]["signatures"][0]) == list(want[0]["signatures"][0])
        assert out.read_text() == sig_path.read_text()  # byte for byte, in fact
    with pytest.raises(ValueError, match="holds no DNA sketch with ksize=21"):
        sigfile.read_sig(sig_path, ksize=21)


# ------------------------------------------------------------------------------------- method wrapper
def test_prepare_genomes_bad_method(tmp_path: Path, input_genomes_tiny: Path) -> None:
    db = tmp_path / "bad-args.db"
    _setup_run(db, input_genomes_tiny, method="guessing", program="guestimator", version="0.0a1")
    logger = setup_logger(None)
    with db_orm.connect_to_db(logger, db) as session:
        run = db_orm.load_run(session, run_id=1)
        with pytest.raises(SystemExit, match="Expected run to be for sourmash, not method guessing"):
            next(sourmash.prepare_genomes(logger, run, tmp_path))


def test_prepare_genomes_bad_kmer(tmp_path: Path, input_genomes_tiny: Path) -> None:
    db = tmp_path / "bad-args.db"
    _setup_run(db, input_genomes_tiny, method="sourmash", program="sourmash", version="0.0a1",
               extra="scaled=" + str(sourmash.SCALED))
    logger = setup_logger(None)
    with db_orm.connect_to_db(logger, db) as session:
        run = db_orm.load_run(session, run_id=1)
        with pytest.raises(SystemExit, match=f"sourmash requires a k-mer size, default is {sourmash.KMER_SIZE}"):
            next(sourmash.prepare_genomes(logger, run, cache=tmp_path))


def test_prepare_genomes_bad_extra(tmp_path: Path, input_genomes_tiny: Path) -> None:
    db = tmp_path / "bad-args.db"
    _setup_run(db, input_genomes_tiny, method="sourmash", program="sourmash", version="0.0a1",
               kmersize=sourmash.KMER_SIZE)
    logger = setup_logger(None)
    with db_orm.connect_to_db(logger, db) as session:
        run = db_orm.load_run(session, run_id=1)
        with pytest.raises(SystemExit, match=f"sourmash requires extra setting, default is scaled={sourmash.SCALED}"):
            next(sourmash.prepare_genomes(logger, run, cache=tmp_path))


def test_prepare_genomes_bad_cache(tmp_path: Path, input_genomes_tiny: Path) -> None:
    db = tmp_path / "bad-args.db"
    _setup_run(db, input_genomes_tiny, method="sourmash", program="sourmash", version="0.0a1",
               kmersize=sourmash.KMER_SIZE, extra="scaled=" + str(sourmash.SCALED))
    logger = setup_logger(None)
    with db_orm.connect_to_db(logger, db) as session:
        run = db_orm.load_run(session, run_id=1)
        with pytest.raises(ValueError, match="Cache directory '/does/not/exist' does not exist"):
            next(sourmash.prepare_genomes(logger, run, cache=Path("/does/not/exist")))


def test_parser_with_bad_branchwater(tmp_path: Path) -> None:
    """Self-vs-self must be one; unexpected pairs are fatal; column order is free; blank lines skipped."""
    mock_csv = tmp_path / "faked.csv"
    mock_csv.write_text(
        "max_containment_ani,query_name,match_name,query_containment_ani\n\n"
        "1.0,AAAAAA,AAAAAA,1.0\n0.9,AAAAAA,BBBBBB,0.85\nNaN,BBBBBB,BBBBBB,NaN\n"
    )
    expected = {("AAAAAA", "AAAAAA"), ("AAAAAA", "BBBBBB"), ("BBBBBB", "AAAAAA"), ("BBBBBB", "BBBBBB")}
    logger = setup_logger(None)
    parser = sourmash.parse_sourmash_manysearch_csv(logger, mock_csv, expected)
    assert next(parser) == ("AAAAAA", "AAAAAA", 1.0, 1.0)
    assert next(parser) == ("AAAAAA", "BBBBBB", 0.85, 0.9)
    with pytest.raises(ValueError, match="Expected sourmash manysearch BBBBBB vs self to be one, not 'NaN'"):
        next(parser)
    parser = sourmash.parse_sourmash_manysearch_csv(logger, mock_csv, {("AAAAAA", "AAAAAA")})
    assert next(parser) == ("AAAAAA", "AAAAAA", 1.0, 1.0)
    with pytest.raises(SystemExit, match=r"Did not expect AAAAAA vs BBBBBB in faked\.csv"):
        next(parser)


def test_parser_with_bad_header(tmp_path: Path) -> None:
    mock_csv = tmp_path / "faked.csv"
    mock_csv.write_text("max_containment_ani,query_name,match_name,subject_containment_ani\n")
    logger = setup_logger(None)
    parser = sourmash.parse_sourmash_manysearch_csv(logger, mock_csv, set())
    with pytest.raises(SystemExit, match="Missing expected fields in sourmash manysearch header, found: "
                       "'max_containment_ani,query_name,match_name,subject_containment_ani'"):
        next(parser)


@pytest.mark.parametrize("name", ["viral_example", "bad_alignments", "bacterial_example"])
def test_parser_on_fixture_manysearch(golden: Path, name: str) -> None:
    """The reference's own manysearch.csv files parse to N^2 tuples, None for the row-less pairs."""
    sig_dir = golden / name / "intermediates" / "sourmash"
    hashes = sorted(p.stem for p in sig_dir.glob("*.sig"))
    logger = setup_logger(None)
    rows = list(sourmash.parse_sourmash_manysearch_csv(
        logger, sig_dir / "manysearch.csv", {(q, s) for q in hashes for s in hashes}))
    assert len(rows) == len(hashes) ** 2
    nulls = [r for r in rows if r[2] is None]
    assert len(nulls) == (2 if name == "bad_alignments" else 0)
    assert all(r[3] == 1.0 for r in rows if r[0] == r[1])


def test_compute_bad_args(tmp_path: Path) -> None:
    tmp_json = tmp_path / "bad args.json"
    logger = setup_logger(None)
    with db_orm.connect_to_db(logger, tmp_path / "bad args.db") as session:
        tool = tools.get_sourmash()
        config = db_orm.Configuration(method="sourmash", program=tool.exe_path.name, version=tool.version,
                                      kmersize=31, extra="scaled=1234")
        run = db_orm.Run(configuration=config)
        with pytest.raises(SystemExit, match=("Missing sourmash signatures directory"
                                              f" '{tmp_path}/sourmash_k=31_scaled=1234' - check cache setting.")):
            private_cli.compute_sourmash(logger, tmp_path, session, run, tmp_json, tmp_path, {}, {},
                                         {"ABCDE": 12345}, "HIJKL", cache=tmp_path)
        config.version = "0.0a1"
        with pytest.raises(SystemExit, match="Run configuration was panib200 0.0a1 but we have panib200"):
            private_cli.compute_sourmash(logger, tmp_path, session, run, tmp_json, tmp_path, {}, {},
                                         {"ABCDE": 12345}, "HIJKL", cache=tmp_path)


def test_compute_tile_bad_args(tmp_path: Path) -> None:
    tool = tools.ExternalToolData(exe_path=Path("sourmash"), version="0.0a1")
    logger = setup_logger(None)
    with pytest.raises(ValueError, match="Given cache directory '/does/not/exist' does not exist"):
        next(sourmash.compute_sourmash_tile(logger, tool, {""}, {""}, Path("/does/not/exist"), tmp_path))
    with pytest.raises(SystemExit, match="Missing sourmash signature file "):
        next(sourmash.compute_sourmash_tile(logger, tool, {"ACBDE"}, {"ABCDE"}, tmp_path, tmp_path))


def test_compute_column_bad_args(tmp_path: Path, input_genomes_tiny: Path) -> None:
    db = tmp_path / "bad.db"
    out = tmp_path / "out.json"
    with pytest.raises(SystemExit, match="does not exist"):
        private_cli.compute_column(database=db, run_id=1, subject="1", json=out, log=Path("-"))
    tool = tools.get_sourmash()
    _setup_run(db, input_genomes_tiny, method="sourmash", program=tool.exe_path.stem, version=tool.version,
               kmersize=31, extra="scaled=300")
    with pytest.raises(SystemExit, match="Did not recognise 'XXX' as an MD5 hash, filename, or column number in run-id 1"):
        private_cli.compute_column(database=db, run_id=1, subject="XXX", json=out, log=Path("-"))
    with pytest.raises(SystemExit, match="Single column should be in range 1 to 3, or for some methods 0 meaning all columns, but not 4"):
        private_cli.compute_column(database=db, run_id=1, subject="4", json=out, log=Path("-"))
    with pytest.raises(SystemExit, match="Missing sourmash signatures directory"):
        private_cli.compute_column(database=db, run_id=1, subject="0", json=out, cache=tmp_path, log=Path("-"))
    with pytest.raises(SystemExit, match="Database has no run-id 2"):
        private_cli.compute_column(database=db, run_id=2, subject="0", json=out, log=Path("-"))


def test_manysearch_number_format() -> None:
    """Rust prints f64 as the shortest round-trip decimal without exponent."""
    assert sourmash._fmt(1.0) == "1.0"  # noqa: SLF001
    assert sourmash._fmt(0.8888888888888888) == "0.8888888888888888"  # noqa: SLF001
    assert sourmash._fmt(0.00001234) == "0.00001234"  # noqa: SLF001
    assert sourmash._fmt(0.014417744916820702) == "0.014417744916820702"  # noqa: SLF001


# ------------------------------------------------------------------------------------- C FASTA ingest
def test_c_fasta_parser_matches_python_iterator(golden: Path, tmp_path: Path) -> None:
    """panib_fasta_to_stream == fasta_bytes_iterator records joined by one N (fixtures + mangled input)."""
    from pyani_plus_b200 import engine

    files = [golden / "viral_example" / "OP073605.fasta", golden / "MIBY01000005.fasta",
             golden / "bacterial_example" / "NC_002696.fasta.gz", golden / "bacterial_example" / "NC_011916.fas.gz"]
    messy = tmp_path / "messy.fasta"
    messy.write_bytes(b"junk\n\n>one desc \t\r\nAC GT\r\n\r\nNN\n>two\n\n>three\nacgt\x0b\n>\nTT")
    empty = tmp_path / "empty.fasta"
    empty.write_bytes(b"no records here\nACGT\n")
    for f in [*files, messy, empty]:
        records = utils.read_fasta_records(f)
        stream, n_records, total, title = utils.read_fasta_stream(f)
        assert n_records == len(records)
        assert total == sum(len(s) for _, s in records)
        assert stream.tobytes() == b"N".join(s for _, s in records)
        assert title == (records[0][0] if records else None)
    stream, n, total, title = engine.fasta_to_stream(b"")
    assert (stream.size, n, total, title) == (0, 0, 0, None)


def test_c_fasta_parser_random_text() -> None:
    """Seeded random FASTA-like text (long and short lines, blanks / tabs / CRs at every offset of the 16-byte
    scan, CRLF, stray '>' inside lines, empty records, no final newline): the C parser, in measuring mode and in
    writing mode, equals the Python iterator of utils.py (the restatement of reference utils.py:40-90)."""
    import ctypes
    import io
    import random

    from pyani_plus_b200 import engine

    lib = engine.load_library()
    rng = random.Random(20261017)
    alphabet = [b"A", b"C", b"G", b"T", b"N", b"a", b"c", b"g", b"t", b" ", b"\t", b"\r", b">", b"\x0b", b"-"]
    weights = [30, 30, 30, 30, 4, 4, 4, 4, 4, 3, 2, 3, 1, 1, 1]
    for case in range(300):
        lines = []
        for _ in range(rng.randrange(0, 12)):
            kind = rng.random()
            if kind < 0.25:
                lines.append(b">" + bytes(rng.choices(b"abc \t\r\x0c", k=rng.randrange(0, 6))))
            elif kind < 0.35:
                lines.append(b"")
            else:
                lines.append(b"".join(rng.choices(alphabet, weights, k=rng.randrange(1, 70))))
        eol = b"\r\n" if case % 5 == 0 else b"\n"
        text = eol.join(lines) + (eol if case % 3 else b"")
        records = list(utils.fasta_bytes_iterator(io.BytesIO(text)))
        want = b"N".join(s for _, s in records)
        stream, n_records, total, title = engine.fasta_to_stream(text)
        assert stream.tobytes() == want, (case, text)
        assert (n_records, total) == (len(records), sum(len(s) for _, s in records)), (case, text)
        assert title == (records[0][0] if records else None), (case, text)
        out4 = (ctypes.c_int64 * 4)()
        assert lib.panib_fasta_to_stream(text, len(text), None, 0, out4) == len(want)  # measuring mode
        if want:  # a destination one byte short is an error, not a truncation
            small = ctypes.create_string_buffer(len(want) - 1 or 1)
            assert lib.panib_fasta_to_stream(text, len(text), small, len(want) - 1, out4) == -2


def test_format_u64_matches_python_str() -> None:
    """panib_format_u64 (the decimal text of .sig files) equals str() for every digit count and separator."""
    import numpy as np

    from pyani_plus_b200 import engine, sigfile

    edge = [0, 1, 9, 10, 11, 99, 100, 101, 999, 1000, 2**32 - 1, 2**32, 2**63, 2**64 - 1, 10**19, 10**19 - 1]
    edge += [10**d for d in range(20)] + [10**d - 1 for d in range(1, 20)]
    rng = np.random.default_rng(7)
    rand = (rng.integers(0, 2**63, 2000, dtype=np.uint64) >> rng.integers(0, 63, 2000).astype(np.uint64)).tolist()
    values = np.array(edge + rand, dtype=np.uint64)
    assert engine.format_u64(values, b",") == ",".join(map(str, values.tolist())).encode()
    assert engine.format_u64(values, b"") == "".join(map(str, values.tolist())).encode()
    assert engine.format_u64(values[:1], b",") == b"0"
    assert engine.format_u64(np.array([], dtype=np.uint64)) == b""
    import hashlib

    assert sigfile.sketch_md5sum(values, 31) == hashlib.md5(("31" + "".join(map(str, values.tolist()))).encode()).hexdigest()  # noqa: S324


def test_matrix_cache_skipped_when_too_large_for_sqlite(tmp_path: Path, monkeypatch: pytest.MonkeyPatch) -> None:
    """The reference caches every run matrix as ONE JSON text (db_orm.py:393-466); SQLite takes at most 10^9
    bytes per text, so very large runs skip the cache and the matrix properties rebuild from the comparisons
    table -- same values as the cached form."""
    import logging

    import numpy as np

    from pyani_plus_b200 import db_orm

    logger = logging.getLogger("test")
    db = tmp_path / "big.sqlite"
    hashes = [f"{i:032x}" for i in range(4)]
    rng = np.random.default_rng(3)
    ident = rng.random((4, 4))
    cov = rng.random((4, 4))
    ident[1, 2] = cov[1, 2] = np.nan
    frames = {}
    for limit in (db_orm.MATRIX_CACHE_MAX_BYTES, 100):  # second pass: 4 x 4 x 20 bytes is already "too large"
        monkeypatch.setattr(db_orm, "MATRIX_CACHE_MAX_BYTES", limit)
        with db_orm.connect_to_db(logger, db if limit > 100 else tmp_path / "big2.sqlite") as session:
            config = db_orm.db_configuration(session, "sourmash", "panib200", "0", kmersize=31, extra="scaled=1000",
                                             create=True)
            for h in hashes:
                db_orm.db_genome(logger, session, tmp_path / f"{h}.fna", h, create=True, stats=(10, b"t", False))
            run = db_orm.add_run(session, config, "x", tmp_path, "Running", "t", None,
                                 {tmp_path / f"{h}.fna": h for h in hashes})
            assert db_orm.insert_comparison_arrays(logger, session, config.configuration_id, hashes, hashes, ident, cov)
            run.cache_comparisons()
            run.status = "Done"
            session.commit()
            assert (run.df_identity is None) == (limit == 100)
            frames[limit] = (run.identities, run.cov_query, run.hadamard)
            assert run.aln_length is not None or limit == 100
    for a, b in zip(frames[100], frames[max(frames)], strict=True):
        assert a is not None and b is not None
        # the JSON cache keeps 10 significant digits (pandas' to_json, as in the reference); the table is exact
        np.testing.assert_allclose(a.to_numpy(), b.to_numpy(), rtol=0, atol=1e-9, equal_nan=True)
        assert list(a.index) == list(b.index) == hashes


def test_bulk_insert_literals_and_null_matrices(tmp_path: Path, monkeypatch: pytest.MonkeyPatch) -> None:
    """insert_comparison_arrays writes the per-call constants as SQL literals (quotes in a uname string are
    escaped, not executed) and records exactly what the per-row path records; the all-null matrices of the
    method (aln_length, sim_errors) are cached with the text pandas would produce."""
    import logging
    import platform
    from collections import namedtuple

    import numpy as np
    import pandas as pd

    from pyani_plus_b200 import db_orm

    logger = logging.getLogger("test")
    uname = namedtuple("uname", "system node release version machine processor")  # noqa: PYI024
    monkeypatch.setattr(platform, "uname", lambda: uname("Li'nux", "n", "6.1'); DROP TABLE genomes; --", "v", 'x"86', ""))
    hashes = [f"{i:032x}" for i in range(5)]
    ident = np.random.default_rng(5).random((5, 5))
    cov = ident * 0.5
    ident[0, 3] = cov[0, 3] = np.nan
    with db_orm.connect_to_db(logger, tmp_path / "lit.sqlite") as session:
        config = db_orm.db_configuration(session, "sourmash", "panib200", "0", kmersize=31, extra="scaled=1000",
                                         create=True)
        for h in hashes:
            db_orm.db_genome(logger, session, tmp_path / f"{h}.fna", h, create=True, stats=(10, b"t", False),
                             commit=False)
        run = db_orm.add_run(session, config, "x", tmp_path, "Running", "t", None,
                             {tmp_path / f"{h}.fna": h for h in hashes})
        assert db_orm.insert_comparison_arrays(logger, session, config.configuration_id, hashes, hashes, ident, cov)
        rows = session.execute(
            "SELECT query_hash, subject_hash, configuration_id, identity, aln_length, sim_errors, cov_query,"
            " cov_subject, uname_system, uname_release, uname_machine FROM comparisons ORDER BY comparison_id"
        ).fetchall()
        assert len(rows) == 25 and session.execute("SELECT COUNT(*) FROM genomes").fetchone()[0] == 5
        for r, (q, s) in zip(rows, ((q, s) for q in range(5) for s in range(5)), strict=True):
            want_i = None if np.isnan(ident[q, s]) else ident[q, s]
            want_c = None if np.isnan(cov[q, s]) else cov[q, s]
            assert r == (hashes[q], hashes[s], config.configuration_id, want_i, None, None, want_c, None,
                         "Li'nux", "6.1'); DROP TABLE genomes; --", 'x"86')
        # a second call is ignored row by row (INSERT OR IGNORE on the unique key)
        assert db_orm.insert_comparison_arrays(logger, session, config.configuration_id, hashes, hashes, ident, cov)
        assert session.execute("SELECT COUNT(*) FROM comparisons").fetchone()[0] == 25
        run.cache_comparisons()
        null_frame = pd.DataFrame(data=np.full((5, 5), np.nan), index=hashes, columns=hashes, dtype=float)
        assert run.df_aln_length == run.df_sim_errors == null_frame.to_json(orient="split")
        assert run.df_identity == pd.DataFrame(data=ident, index=hashes, columns=hashes).to_json(orient="split")


@pytest.mark.parametrize(("n_q", "n_s", "per_stmt", "per_commit"), [(13, 13, 5, 40), (3, 200, 128, 2_000_000),
                                                                    (7, 1, 4, 2), (1, 1, 128, 1)])
def test_bulk_insert_statement_and_commit_boundaries(  # noqa: PLR0913
    tmp_path: Path, n_q: int, n_s: int, per_stmt: int, per_commit: int,
) -> None:
    """Rows are packed into multi-row statements across query-row and commit-chunk boundaries: every ordered
    pair arrives exactly once, in order, with its own values, whatever the block shape."""
    import logging

    import numpy as np

    from pyani_plus_b200 import db_orm

    logger = logging.getLogger("test")
    queries = [f"{i:032x}" for i in range(n_q)]
    subjects = [f"{i + 1000:032x}" for i in range(n_s)]
    rng = np.random.default_rng(n_q * 1000 + n_s)
    ident = rng.random((n_q, n_s))
    cov = rng.random((n_q, n_s))
    ident[rng.random((n_q, n_s)) < 0.2] = np.nan
    cov[np.isnan(ident)] = np.nan
    with db_orm.connect_to_db(logger, tmp_path / "b.sqlite") as session:
        config = db_orm.db_configuration(session, "sourmash", "panib200", "0", kmersize=31, extra="scaled=1000",
                                         create=True)
        for h in [*queries, *subjects]:
            db_orm.db_genome(logger, session, tmp_path / f"{h}.fna", h, create=True, stats=(10, b"t", False),
                             commit=False)
        session.commit()
        assert db_orm.insert_comparison_arrays(logger, session, config.configuration_id, queries, subjects, ident,
                                               cov, rows_per_commit=per_commit, rows_per_statement=per_stmt)
        rows = session.execute(
            "SELECT query_hash, subject_hash, identity, cov_query FROM comparisons ORDER BY comparison_id").fetchall()
    want = [(queries[q], subjects[s], None if np.isnan(ident[q, s]) else ident[q, s],
             None if np.isnan(cov[q, s]) else cov[q, s]) for q in range(n_q) for s in range(n_s)]
    assert rows == want


def test_missing_block_counts_in_sql(tmp_path: Path) -> None:
    """``missing_block`` (resume: reference private_cli.py:879-898 computes only what a column lacks) equals the
    row-by-row definition on partially filled runs, incl. comparisons that belong to genomes outside the run."""
    import logging

    import numpy as np

    from pyani_plus_b200 import db_orm, private_cli

    logger = logging.getLogger("test")
    rng = np.random.default_rng(11)
    hashes = [f"{i:032x}" for i in range(9)]
    outsider = "f" * 32
    for case in range(6):
        with db_orm.connect_to_db(logger, tmp_path / f"m{case}.sqlite") as session:
            config = db_orm.db_configuration(session, "sourmash", "panib200", "0", kmersize=31,
                                             extra="scaled=1000", create=True)
            for h in [*hashes, outsider]:
                db_orm.db_genome(logger, session, tmp_path / f"{h}.fna", h, create=True, stats=(10, b"t", False),
                                 commit=False)
            run = db_orm.add_run(session, config, "x", tmp_path, "Running", "t", None,
                                 {tmp_path / f"{h}.fna": h for h in hashes})
            keep = rng.random((9, 9)) < (0.0, 0.5, 0.9, 1.0, 0.97, 0.2)[case]
            if case == 4:
                keep[:, 3] = False  # one subject without any comparison
            for q in range(9):
                present = [hashes[s] for s in range(9) if keep[q, s]]
                if present:
                    vals = np.full(len(present), 0.5)
                    db_orm.insert_comparison_arrays(logger, session, config.configuration_id, [hashes[q]], present,
                                                    vals, vals)
            # rows of a genome that is not part of the run must not count
            db_orm.insert_comparison_arrays(logger, session, config.configuration_id, [outsider], hashes,
                                            np.full(9, 0.1), np.full(9, 0.1))
            queries, subjects = private_cli.missing_block(run)
            want_s = {hashes[s] for s in range(9) if not keep[:, s].all()}
            want_q = {hashes[q] for s in range(9) for q in range(9) if not keep[q, s]}
            assert (queries, subjects) == (want_q, want_s), case


def test_md5_failure_modes_and_logger_flag(golden: Path, tmp_path: Path) -> None:
    """Reference tests/test_utils.py: test_md5_str / test_md5_path :36-49 (gzip, str and Path), test_md5_invalid
    :52-63 (missing file, broken symlink), test_setup_logger_dynamic :103-110; plus ``fasta_file_stats``, the
    one-pass form the CLI uses, which must fail the same way."""
    from pyani_plus_b200 import setup_logger

    gz = golden / "bacterial_example" / "NC_002696.fasta.gz"
    assert utils.file_md5sum(str(gz)) == utils.file_md5sum(gz) == "f19cb07198a41a4406a22b2f57a6b5e7"
    md5, total, title, is_gzip = utils.fasta_file_stats(gz)
    assert (md5, total, is_gzip) == ("f19cb07198a41a4406a22b2f57a6b5e7", 4016947, True)
    assert title is not None and title.startswith(b"NC_002696")
    bad_link = tmp_path / "bad-link.fasta"
    bad_link.symlink_to("/does/not/exist.fasta")
    for fn in (utils.file_md5sum, utils.fasta_file_stats):
        with pytest.raises(ValueError, match=r"Input /does/not/exist\.txt not found"):
            fn("/does/not/exist.txt")
        with pytest.raises(ValueError, match=r"Input .*/bad-link\.fasta is a broken symlink"):
            fn(bad_link)
    with pytest.raises(SystemExit, match="ERROR: Internal flag value for dynamic log setting unresolved"):
        setup_logger(Path("--"))
