"""``import-comparisons``: the JSON hand-over between a worker and the database, as the reference's
tests/test_json.py exercises it (test_json_import_errors_core :36-98, test_json_import_errors :101-230)."""

from __future__ import annotations

from pathlib import Path

import pytest

from pyani_plus_b200 import db_orm, private_cli, setup_logger

GUESS = '"configuration":{"method":"guessing", "program":"guestimate","version":"0.1.2beta3", "fragsize":100,"kmersize":51}'
UNAME = '"uname":{"system":"Darwin", "release":"24.3.0", "machine":"arm64"}'


def _log_guess_config(tmp_db: Path) -> None:
    private_cli.log_configuration(tmp_db, method="guessing", program="guestimate", version="0.1.2beta3",
                                  fragsize=100, kmersize=51, create_db=True)


def _log_two_genomes(tmp_db: Path, input_genomes_tiny: Path) -> None:
    private_cli.log_genome(database=tmp_db, fasta=[input_genomes_tiny / "MGV-GENOME-0264574.fas",
                                                   input_genomes_tiny / "MGV-GENOME-0266457.fna"])


def test_json_import_errors_core(input_genomes_tiny: Path, tmp_path: Path) -> None:
    tmp_db = tmp_path / "json.sqlite"
    tmp_json = tmp_path / "x.json"
    tmp_json.touch()
    with pytest.raises(SystemExit, match=f"Database '{tmp_db}' does not exist"):
        private_cli.import_comparisons(tmp_db, json=[tmp_json], debug=False, log=Path("-"))
    tmp_db.touch()
    with pytest.raises(SystemExit, match="does not contain any configurations"):
        private_cli.import_comparisons(tmp_db, json=[tmp_json], debug=False, log=Path("-"))
    _log_guess_config(tmp_db)
    with pytest.raises(SystemExit, match="does not contain any genomes"):
        private_cli.import_comparisons(tmp_db, json=[tmp_json], debug=False, log=Path("-"))
    _log_two_genomes(tmp_db, input_genomes_tiny)
    tmp_json.write_text("[")
    with pytest.raises(SystemExit, match=f"JSON file '{tmp_json}' invalid"):
        private_cli.import_comparisons(tmp_db, json=[tmp_json], debug=False, log=Path("-"))
    tmp_json.write_text("[]")
    with pytest.raises(SystemExit, match=f"JSON file '{tmp_json}' does not use the expected structure"):
        private_cli.import_comparisons(tmp_db, json=[tmp_json], debug=False, log=Path("-"))


def test_json_import_errors(caplog: pytest.LogCaptureFixture, input_genomes_tiny: Path, tmp_path: Path) -> None:
    tmp_db = tmp_path / "json.sqlite"
    tmp_json = tmp_path / "x.json"
    _log_guess_config(tmp_db)
    _log_two_genomes(tmp_db, input_genomes_tiny)

    tmp_json.touch()  # an empty file: a worker that was interrupted before it wrote anything
    caplog.clear()
    private_cli.import_comparisons(tmp_db, json=[tmp_json], debug=True, log=Path("-"))
    assert f"JSON file '{tmp_json}' is empty" in caplog.text
    assert f"Imported 0 from '{tmp_json}'" in caplog.text

    anim = '"configuration":{"method":"ANIm", "program":"nucmer","version":"3.1", "mode":"mum"}'
    for text, message in (
        ("{" + anim + ", " + UNAME + ', "comparisons":[]}', "configuration not in database"),
        ("{" + anim + ', "uname":{"system":"Darwin", "release":"24.3.0"}, "comparisons":[]}', "uname incomplete"),
        ('{"configuration":{"method":"ANIm"}, ' + UNAME + ', "comparisons":[]}', "configuration incomplete"),
        ("{" + GUESS + ", " + UNAME + ', "comparisons":[{"query_hash":"689d3fd6881db36b5e08329cf23cecdd", '
         '"identity":0.99}]}', r"comparison\(s\) incomplete"),
    ):
        tmp_json.write_text(text)
        with pytest.raises(SystemExit, match=f"JSON file '{tmp_json}' {message}"):
            private_cli.import_comparisons(tmp_db, json=[tmp_json], debug=False, log=Path("-"))

    tmp_json.write_text("{" + GUESS + ", " + UNAME + ', "comparisons":[]}')  # a warning only
    caplog.clear()
    private_cli.import_comparisons(tmp_db, json=[tmp_json], debug=False, log=Path("-"))
    assert f"JSON file '{tmp_json}' has no comparisons" in caplog.text

    tmp_json.write_text("{" + GUESS + ", " + UNAME + ', "comparisons":[{"query_hash":"689d3fd6881db36b5e08329cf23cecdd",'
                        ' "subject_hash":"78975d5144a1cd12e98898d573cf6536", "identity":0.99}]}')
    private_cli.import_comparisons(tmp_db, json=[tmp_json], debug=False, log=Path("-"))
    private_cli.import_comparisons(tmp_db, json=[tmp_json, tmp_json], debug=False, log=Path("-"))  # INSERT OR IGNORE
    with db_orm.connect_to_db(setup_logger(None), tmp_db) as session:
        rows = session.execute("SELECT query_hash, subject_hash, identity, cov_query, uname_system, uname_release,"
                               " uname_machine FROM comparisons").fetchall()
    assert rows == [("689d3fd6881db36b5e08329cf23cecdd", "78975d5144a1cd12e98898d573cf6536", 0.99, None, "Darwin",
                     "24.3.0", "arm64")]
