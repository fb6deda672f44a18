"""``pyani-plus resume`` on the host side: everything that is decided before any GPU work.

Mirrors the reference's tests/test_public_cli.py scenarios for resume (test_resume_empty :179-193,
test_resume_dir_gone :1580-1619, test_resume_unknown :1622-1659, test_resume_complete :1662-1724,
test_resume_fasta_gone :1727-1810) with the sourmash method recorded as this engine records it.
"""

from __future__ import annotations

import logging
from pathlib import Path

import pytest

from pyani_plus_b200 import db_orm, public_cli, setup_logger, tools
from pyani_plus_b200.utils import file_md5sum


def _record_run(  # noqa: PLR0913
    tmp_db: Path, fasta_dir: Path, genomes_from: Path, *, method: str = "sourmash", status: str = "Partial",
    complete: bool = False, program: str | None = None, version: str | None = None,
) -> dict[Path, str]:
    tool = tools.get_sourmash()
    logger = setup_logger(None)
    with db_orm.connect_to_db(logger, tmp_db) as session:
        config = db_orm.db_configuration(session, method, program or tool.exe_path.stem, version or tool.version,
                                         kmersize=31, extra="scaled=300", create=True)
        fasta_to_hash = {f: file_md5sum(f) for f in sorted(genomes_from.glob("*.f*"))}
        for filename, md5 in fasta_to_hash.items():
            db_orm.db_genome(logger, session, filename, md5, create=True)
        if complete:
            for q in fasta_to_hash.values():
                for s in fasta_to_hash.values():
                    db_orm.db_comparison(session, config.configuration_id, q, s, 1.0 if q == s else 0.99, 12345)
        db_orm.add_run(session, config, cmdline="pyani-plus sourmash ...", fasta_directory=fasta_dir, status=status,
                       name="Testing resume", fasta_to_hash=fasta_to_hash)
    return fasta_to_hash


def test_resume_empty(tmp_path: Path) -> None:
    with pytest.raises(SystemExit, match="Database /does/not/exist does not exist"):
        public_cli.resume(database=Path("/does/not/exist"))
    tmp_db = tmp_path / "resume-empty.sqlite"
    with db_orm.connect_to_db(setup_logger(None), tmp_db):
        pass
    with pytest.raises(SystemExit, match=r"Database contains no runs\."):
        public_cli.resume(database=tmp_db)
    with pytest.raises(SystemExit, match=r"Database has no run-id 1\."):
        public_cli.resume(database=tmp_db, run_id=1)


def test_resume_dir_gone(tmp_path: Path, input_genomes_tiny: Path) -> None:
    tmp_db = tmp_path / "resume.sqlite"
    _record_run(tmp_db, Path("/mnt/shared/old"), input_genomes_tiny)
    with pytest.raises(SystemExit,
                       match=r"run-id 1 used input folder /mnt/shared/old, but that is not a directory \(now\)."):
        public_cli.resume(database=tmp_db)


def test_resume_unknown_method(tmp_path: Path, input_genomes_tiny: Path) -> None:
    tmp_db = tmp_path / "resume.sqlite"
    _record_run(tmp_db, input_genomes_tiny, input_genomes_tiny, method="guessing")
    with pytest.raises(SystemExit, match=r"Unknown method guessing for run-id 1 in .*/resume\.sqlite"):
        public_cli.resume(database=tmp_db)


def test_resume_other_engine_version(tmp_path: Path, input_genomes_tiny: Path) -> None:
    """A run recorded by real sourmash (or another version of this engine) is never continued by this one
    (reference public_cli.py:773-785: same check on tool name and version)."""
    tmp_db = tmp_path / "resume.sqlite"
    _record_run(tmp_db, input_genomes_tiny, input_genomes_tiny, program="sourmash", version="4.8.11")
    with pytest.raises(SystemExit, match=r"We have panib200 version .*, but run-id 1 used sourmash version 4\.8\.11 instead\."):
        public_cli.resume(database=tmp_db)


def test_resume_complete(caplog: pytest.LogCaptureFixture, tmp_path: Path, input_genomes_tiny: Path) -> None:
    """A complete run is recognised from the database alone and left as it is (no GPU is touched)."""
    caplog.set_level(logging.INFO)
    tmp_db = tmp_path / "resume.sqlite"
    _record_run(tmp_db, input_genomes_tiny, input_genomes_tiny, status="Complete", complete=True)
    assert public_cli.resume(database=tmp_db) == 0
    assert "Resuming run-id 1\n" in caplog.text
    assert "Database already has all 3²=9 sourmash comparisons" in caplog.text
    with db_orm.connect_to_db(setup_logger(None), tmp_db) as session:
        assert [c.identity for c in session.get_run(1).comparisons()].count(0.99) == 6


def test_resume_fasta_gone(caplog: pytest.LogCaptureFixture, tmp_path: Path, input_genomes_tiny: Path) -> None:
    caplog.set_level(logging.INFO)
    indir = tmp_path / "input"
    indir.mkdir()
    tmp_db = tmp_path / "resume.sqlite"
    fasta_to_hash = _record_run(tmp_db, indir, input_genomes_tiny, status="Complete", complete=True)
    for filename in list(fasta_to_hash)[:-1]:
        (indir / filename.name).symlink_to(filename)
    missing = list(fasta_to_hash)[-1]
    with pytest.raises(SystemExit, match=(f"run-id 1 used .*/{missing.name} with MD5 {fasta_to_hash[missing]}"
                                          " but this FASTA file no longer exists")):
        public_cli.resume(database=tmp_db)
    # with every file present it works, although the files now live in another directory than the one the
    # genomes table remembers (as could happen via an older run)
    (indir / missing.name).symlink_to(missing)
    caplog.clear()
    assert public_cli.resume(database=tmp_db) == 0
    assert "Database already has all 3²=9 sourmash comparisons" in caplog.text
