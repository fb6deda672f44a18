"""``pyani-plus list-runs`` / ``export-run`` on databases built without a GPU.

Mirrors the reference's tests/test_public_cli.py: test_list_runs_empty :164-176, test_export_run_failures
:371-419, test_export_duplicate_stem :422-483 (export-run part; plot-run and classify are out of scope).
"""

from __future__ import annotations

from pathlib import Path

import pytest

from pyani_plus_b200 import db_orm, public_cli, setup_logger
from pyani_plus_b200.utils import file_md5sum


def test_list_runs_empty(capsys: pytest.CaptureFixture[str], tmp_path: Path) -> None:
    with pytest.raises(SystemExit, match="Database /does/not/exist does not exist"):
        public_cli.list_runs(database=Path("/does/not/exist"))
    tmp_db = tmp_path / "list runs empty.sqlite"
    with db_orm.connect_to_db(setup_logger(None), tmp_db):
        pass
    public_cli.list_runs(database=tmp_db)
    assert " 0 analysis runs in " in capsys.readouterr().out


def test_export_run_failures(tmp_path: Path) -> None:
    with pytest.raises(SystemExit, match="Database /does/not/exist does not exist"):
        public_cli.export_run(database=Path("/does/not/exist"), outdir=tmp_path)
    tmp_db = tmp_path / "empty.sqlite"
    tmp_db.touch()
    with pytest.raises(SystemExit, match=r"Database contains no runs\."):
        public_cli.export_run(database=tmp_db, outdir=tmp_path)
    tmp_db = tmp_path / "export.sqlite"
    with db_orm.connect_to_db(setup_logger(None), tmp_db) as session:
        config = db_orm.db_configuration(session, "sourmash", "panib200", "1.2.3", kmersize=31, extra="scaled=300",
                                         create=True)
        for name in ("Trial A", "Trial B"):
            db_orm.add_run(session, config, cmdline="pyani-plus sourmash ...", fasta_directory=Path("/does/not/exist"),
                           status="Empty", name=name)
    with pytest.raises(SystemExit, match=r"Database has no run-id 3\."):
        public_cli.export_run(database=tmp_db, outdir=tmp_path, run_id=3)
    with pytest.raises(SystemExit, match="run-id 1 has no comparisons"):
        public_cli.export_run(database=tmp_db, outdir=tmp_path, run_id=1)
    with pytest.raises(SystemExit, match="run-id 2 has no comparisons"):  # defaults to the latest run
        public_cli.export_run(database=tmp_db, outdir=tmp_path)


def _mock_run(tmp_db: Path, fasta_dir: Path, files: list[Path], *, with_null: bool = False) -> dict[Path, str]:
    logger = setup_logger(None)
    with db_orm.connect_to_db(logger, tmp_db) as session:
        config = db_orm.db_configuration(session, "sourmash", "panib200", "1.2.3", kmersize=31, extra="scaled=300",
                                         create=True)
        fasta_to_hash = {f: file_md5sum(f) for f in files}
        for fasta, md5 in fasta_to_hash.items():
            db_orm.db_genome(logger, session, fasta, md5, create=True)
        db_orm.add_run(session, config, cmdline="pyani-plus sourmash ...", fasta_directory=fasta_dir, status="Done",
                       name="Trial B", fasta_to_hash=fasta_to_hash)
        hashes = sorted(fasta_to_hash.values())
        for q in hashes:
            for s in hashes:
                if with_null and (q, s) == (hashes[0], hashes[-1]):
                    db_orm.db_comparison(session, config.configuration_id, q, s, None, None)
                else:
                    same = q == s
                    db_orm.db_comparison(session, config.configuration_id, q, s, 1.0 if same else 0.99, None,
                                         cov_query=1.0 if same else 0.95)
        session.commit()
    return fasta_to_hash


def test_export_duplicate_stem(tmp_path: Path, input_genomes_tiny: Path) -> None:
    fasta = tmp_path / "genomes"
    fasta.mkdir()
    (fasta / "example.fasta").symlink_to(input_genomes_tiny / "OP073605.fasta")
    (fasta / "example.fna").symlink_to(input_genomes_tiny / "MGV-GENOME-0266457.fna")
    (fasta / "example.fas").symlink_to(input_genomes_tiny / "MGV-GENOME-0264574.fas")
    tmp_db = tmp_path / "dup-stems.db"
    _mock_run(tmp_db, input_genomes_tiny, sorted(fasta.glob("*.fa*")))
    with pytest.raises(SystemExit, match=r"Duplicate filename stems, consider using MD5 labelling\."):
        public_cli.export_run(database=tmp_db, outdir=tmp_path / "out1")
    assert public_cli.export_run(database=tmp_db, outdir=tmp_path / "out2", label="md5") == 0
    assert public_cli.export_run(database=tmp_db, outdir=tmp_path / "out3", label="filename") == 0
    head = (tmp_path / "out3" / "sourmash_identity.tsv").read_text().splitlines()[0]
    assert head.split("\t")[1:] == ["example.fas", "example.fasta"]  # "*.fa*" leaves the .fna out, as in the reference test


def test_list_and_export_count_nulls_in_sql(capsys: pytest.CaptureFixture[str], tmp_path: Path,
                                            input_genomes_tiny: Path) -> None:
    """Done / Null / Miss come from SQL counts and the long-form table is streamed in (query, subject) order."""
    tmp_db = tmp_path / "nulls.db"
    files = sorted(input_genomes_tiny.glob("*.f*"))
    fasta_to_hash = _mock_run(tmp_db, input_genomes_tiny, files, with_null=True)
    public_cli.list_runs(database=tmp_db)
    out = capsys.readouterr().out
    assert " 1 analysis runs in " in out
    row = next(line for line in out.splitlines() if "sourmash" in line and "Trial" in line)
    cells = [c.strip() for c in row.replace("┃", "│").split("│")]
    assert cells[4:8] == ["8", "1", "0", "9=3²"], cells
    assert public_cli.export_run(database=tmp_db, outdir=tmp_path / "out", label="filename") == 0
    lines = (tmp_path / "out" / "sourmash_run_1.tsv").read_text().splitlines()
    assert lines[0] == "#Query\tSubject\tIdentity\tQuery-Cov\tSubject-Cov\tHadamard\ttANI\tAlign-Len\tSim-Errors"
    assert len(lines) == 10  # noqa: PLR2004
    by_hash = {md5: f.name for f, md5 in fasta_to_hash.items()}
    hashes = sorted(by_hash)
    want = [(by_hash[q], by_hash[s]) for q in hashes for s in hashes]
    assert [tuple(line.split("\t")[:2]) for line in lines[1:]] == want
    assert lines[3] == f"{by_hash[hashes[0]]}\t{by_hash[hashes[2]]}" + "\tNA" * 7  # the NULL pair
    assert lines[1] == f"{by_hash[hashes[0]]}\t{by_hash[hashes[0]]}\t1.0\t1.0\tNA\t1.0\t-0.0\tNA\tNA"
    assert lines[2].split("\t")[2:7] == ["0.99", "0.95", "NA", str(0.99 * 0.95), str(-__import__("math").log(0.99 * 0.95))]
    for name in ("identity", "aln_lengths", "sim_errors", "query_cov", "hadamard", "tANI"):
        assert (tmp_path / "out" / f"sourmash_{name}.tsv").is_file()


def test_partial_run(caplog: pytest.LogCaptureFixture, capsys: pytest.CaptureFixture[str], tmp_path: Path,
                     input_genomes_tiny: Path) -> None:
    """list-runs and export-run on mock data with an empty, a partial and a complete run (reference
    tests/test_public_cli.py:196-330; its delete-run part is out of scope)."""
    import logging

    caplog.set_level(logging.INFO)
    tmp_db = tmp_path / "list runs.sqlite"
    logger = setup_logger(None)
    with db_orm.connect_to_db(logger, tmp_db) as session:
        config = db_orm.db_configuration(session, "sourmash", "sourmash", "1.2.3", kmersize=31, extra="scaled=300",
                                         create=True)
        fasta_to_hash = {f: file_md5sum(f) for f in sorted(input_genomes_tiny.glob("*.f*"))}
        for filename, md5 in fasta_to_hash.items():
            db_orm.db_genome(logger, session, filename, md5, create=True)
        for q in list(fasta_to_hash.values())[1:]:  # 4 of the possible 9 comparisons
            for s in list(fasta_to_hash.values())[1:]:
                db_orm.db_comparison(session, config.configuration_id, q, s, 1.0 if q == s else 0.99, 12345)
        common = {"cmdline": "pyani-plus sourmash ...", "fasta_directory": input_genomes_tiny}
        db_orm.add_run(session, config, status="Empty", name="Trial A", fasta_to_hash={}, **common)
        db_orm.add_run(session, config, status="Running", name="Trial B", fasta_to_hash=fasta_to_hash, **common)
        db_orm.add_run(session, config, status="Done", name="Trial C",
                       fasta_to_hash=dict(list(fasta_to_hash.items())[1:]), **common)
    public_cli.list_runs(database=tmp_db)
    output = capsys.readouterr().out
    assert " 3 analysis runs in " in output, output
    assert " Method   ┃ Done ┃ Null ┃ Miss ┃ Total ┃ Status " in output, output
    assert " sourmash │    0 │    0 │    0 │  0=0² │ Empty " in output, output
    assert " sourmash │    4 │    0 │    5 │  9=3² │ Running " in output, output
    assert " sourmash │    4 │    0 │    0 │  4=2² │ Done " in output, output

    # run 3 is complete although .cache_comparisons() has not happened yet: export-run does it
    for label, header in (("md5", "\t5584c7029328dc48d33f95f0a78f7e57\t78975d5144a1cd12e98898d573cf6536\n"),
                          ("stem", "\tMGV-GENOME-0266457\tOP073605\n"),
                          ("filename", "\tMGV-GENOME-0266457.fna\tOP073605.fasta\n")):
        caplog.clear()
        public_cli.export_run(database=tmp_db, run_id=3, outdir=tmp_path, label=label)
        assert f"Wrote matrices to {tmp_path}" in caplog.text
        with (tmp_path / "sourmash_identity.tsv").open() as handle:
            assert handle.readline() == header

    # run 2 is partial: the long-form table is written, the matrices are refused
    with pytest.raises(SystemExit, match="run-id 2 has only 4 of 3²=9 comparisons, 5 needed"):
        public_cli.export_run(database=tmp_db, run_id=2, outdir=tmp_path)
    with (tmp_path / "sourmash_run_2.tsv").open() as handle:
        assert handle.readline().rstrip("\n").split("\t") == [
            "#Query", "Subject", "Identity", "Query-Cov", "Subject-Cov", "Hadamard", "tANI", "Align-Len", "Sim-Errors"]
        assert sum(1 for _ in handle) == 4  # noqa: PLR2004

    with pytest.raises(SystemExit, match=r"We have panib200 version .*, but run-id 2 used sourmash version 1\.2\.3 instead\."):
        public_cli.resume(database=tmp_db, run_id=2)
    with pytest.raises(SystemExit, match=r"No genomes recorded for run-id 1, cannot resume\."):
        public_cli.resume(database=tmp_db, run_id=1)
