"""Low-level commands of ``.pyani-plus-private-cli`` that never touch the GPU.

Mirrors the reference's tests/test_private_cli.py: test_log_configuration :38-80, test_log_genome :83-103,
test_log_run :106-172, test_log_comparison_serial_and_skip_process_genomes :278-359 (the comparisons are
recorded through ``db_orm.db_comparison``: this repo has no ``log-comparison`` command), test_missing_db
:460-474, test_prepare_genomes_bad_args :477-499.
"""

from __future__ import annotations

from pathlib import Path

import pytest

from pyani_plus_b200 import db_orm, private_cli, setup_logger
from pyani_plus_b200.utils import file_md5sum


def test_log_configuration(caplog: pytest.LogCaptureFixture, tmp_path: Path) -> None:
    tmp_db = tmp_path / "new.sqlite"
    with pytest.raises(SystemExit, match="does not exist, but not using --create-db"):
        private_cli.log_configuration(tmp_db, method="guessing", program="guestimate", version="0.1.2beta3",
                                      fragsize=100, kmersize=51, create_db=False)
    private_cli.log_configuration(tmp_db, method="guessing", program="guestimate", version="0.1.2beta3",
                                  fragsize=100, kmersize=51, create_db=True)
    assert "Configuration identifier 1" in caplog.text
    caplog.clear()
    private_cli.log_configuration(tmp_db, method="guessing", program="guestimate", version="0.1.2beta3",
                                  fragsize=75, kmersize=31, create_db=False)
    assert "Configuration identifier 2" in caplog.text
    caplog.clear()  # the same settings again: the existing entry is returned, not duplicated
    private_cli.log_configuration(tmp_db, method="guessing", program="guestimate", version="0.1.2beta3",
                                  fragsize=75, kmersize=31, create_db=False)
    assert "Configuration identifier 2" in caplog.text


def test_log_genome(tmp_path: Path, input_genomes_tiny: Path) -> None:
    tmp_db = tmp_path / "new.sqlite"
    subset = sorted(input_genomes_tiny.glob("*.fasta"))
    with pytest.raises(SystemExit, match="does not exist, but not using --create-db"):
        private_cli.log_genome(database=tmp_db, fasta=subset)
    private_cli.log_genome(database=tmp_db, fasta=subset, create_db=True)
    with db_orm.connect_to_db(setup_logger(None), tmp_db) as session:
        for filename in subset:
            genome = session.get_genome(file_md5sum(filename))
            assert genome is not None and genome.path == str(filename) and genome.length > 0


def test_log_run(caplog: pytest.LogCaptureFixture, tmp_path: Path) -> None:
    tmp_db = tmp_path / "new.sqlite"
    common = {"cmdline": "pyani_plus run ...", "name": "Guess Run", "status": "Completed", "method": "guessing",
              "program": "guestimate", "version": "0.1.2beta3", "fragsize": 100, "kmersize": 51}
    with pytest.raises(SystemExit, match="does not exist, but not using --create-db"):
        private_cli.log_run(database=tmp_db, fasta=Path("/does/not/exist/"), create_db=False, **common)
    with pytest.raises(SystemExit, match="No FASTA input genomes under"):
        private_cli.log_run(database=tmp_db, fasta=tmp_path, create_db=True, **common)
    (tmp_path / "example.fasta").write_text(">Tiny\nACGTACGTTA\n")
    caplog.clear()
    private_cli.log_run(database=tmp_db, fasta=tmp_path, create_db=True, **common)
    assert "Run identifier 1" in caplog.text
    with db_orm.connect_to_db(setup_logger(None), tmp_db) as session:
        run = session.get_run(1)
        (link,) = run.fasta_hashes
        assert (run.status, run.name, link.fasta_filename) == ("Completed", "Guess Run", "example.fasta")
        assert run.genomes.one().length == 10 and run.genomes.one().description == "Tiny"  # noqa: PLR2004


def test_complete_run_skips_preparation(caplog: pytest.LogCaptureFixture, tmp_path: Path,
                                        input_genomes_tiny: Path) -> None:
    """A mock database built step by step; prepare-genomes sees that nothing is left to compute."""
    tmp_db = tmp_path / "serial.sqlite"
    settings = {"method": "sourmash", "program": "guestimate", "version": "0.1.2beta3", "kmersize": 51,
                "extra": "scaled=1234"}
    private_cli.log_configuration(tmp_db, create_db=True, **settings)
    assert "Configuration identifier 1" in caplog.text
    fasta = sorted(input_genomes_tiny.glob("*.f*"))
    private_cli.log_genome(database=tmp_db, fasta=fasta)
    with db_orm.connect_to_db(setup_logger(None), tmp_db) as session:
        for query in fasta:
            for subject in fasta:
                db_orm.db_comparison(session, 1, file_md5sum(query), file_md5sum(subject),
                                     1.0 if query == subject else 0.96, 12345, 1, 0.98, 0.98)
        session.commit()
    caplog.clear()
    private_cli.log_run(database=tmp_db, cmdline="pyani_plus run ...", name="Guess Run", status="Completed",
                        fasta=input_genomes_tiny, create_db=False, **settings)
    assert "Run identifier 1" in caplog.text
    with db_orm.connect_to_db(setup_logger(None), tmp_db) as session:
        run = session.get_run(1)
        assert run.comparisons().count() == len(fasta) ** 2
        assert run.identities is not None and float(run.identities.to_numpy().min()) == 0.96  # noqa: PLR2004
    caplog.clear()
    private_cli.prepare_genomes(database=tmp_db, run_id=1, cache=tmp_path)
    assert "Skipping preparation, run already has all 9=3² pairwise values" in caplog.text


def test_missing_db(tmp_path: Path) -> None:
    tmp_db = tmp_path / "new.sqlite"
    with pytest.raises(SystemExit, match="does not exist"):
        private_cli.prepare_genomes(database=tmp_db, run_id=1)
    with pytest.raises(SystemExit, match="does not exist"):
        private_cli.compute_column(database=tmp_db, run_id=1, subject="1", json=tmp_path / "out.json", log=Path("-"))


def test_prepare_genomes_unknown_method(tmp_path: Path, input_genomes_tiny: Path) -> None:
    tmp_db = tmp_path / "bad.sqlite"
    private_cli.log_run(fasta=input_genomes_tiny, database=tmp_db, cmdline="pyani-plus sourmash ...",
                        status="Testing", name="Testing compute-column", method="guessing", program="guestimate",
                        version="0.1.2beta3", create_db=True)
    with pytest.raises(SystemExit, match=r"Unknown method guessing, check tool version\?"):
        private_cli.prepare_genomes(database=tmp_db, run_id=1, cache=tmp_path)
