"""The typer command-line front ends parse the reference's options (no GPU needed: errors come first)."""

from __future__ import annotations

from pathlib import Path

from typer.testing import CliRunner

from pyani_plus_b200 import private_cli, public_cli

runner = CliRunner()


def test_public_help_lists_the_sourmash_commands() -> None:
    result = runner.invoke(public_cli.app, ["--help"])
    assert result.exit_code == 0
    for command in ("sourmash", "resume", "list-runs", "export-run"):
        assert command in result.stdout
    result = runner.invoke(public_cli.app, ["sourmash", "--help"])
    assert result.exit_code == 0
    for option in ("--database", "--name", "--create-db", "--executor", "--cache", "--temp", "--wtemp", "--log",
                   "--scaled", "--kmersize", "--debug"):
        assert option in result.stdout, option
    assert "1000" in result.stdout and "31" in result.stdout  # the reference's defaults
    result = runner.invoke(public_cli.app, ["--version"])
    assert result.exit_code == 0 and "pyANI-plus" in result.stdout


def test_public_argument_errors(tmp_path: Path, input_genomes_tiny: Path) -> None:
    result = runner.invoke(public_cli.app, ["sourmash", str(input_genomes_tiny), "-d", str(tmp_path / "x.db")])
    assert result.exit_code == 1
    assert "does not exist, but not using --create-db" in str(result.exception)
    result = runner.invoke(public_cli.app, ["sourmash", str(tmp_path / "nope"), "-d", str(tmp_path / "x.db"),
                                            "--create-db"])
    assert result.exit_code == 1 and "is not a directory" in str(result.exception)
    result = runner.invoke(public_cli.app, ["sourmash", str(input_genomes_tiny), "-d", str(tmp_path / "x.db"),
                                            "--create-db", "--scaled", "0"])
    assert result.exit_code == 2  # typer range check: scaled >= 1
    result = runner.invoke(public_cli.app, ["resume", "-d", str(tmp_path / "missing.db")])
    assert result.exit_code == 1 and "does not exist" in str(result.exception)
    result = runner.invoke(public_cli.app, ["list-runs", "-d", str(tmp_path / "missing.db")])
    assert result.exit_code == 1


def test_private_help_and_errors(tmp_path: Path) -> None:
    result = runner.invoke(private_cli.app, ["--help"])
    assert result.exit_code == 0
    for command in ("prepare-genomes", "compute-column", "log-run", "log-configuration", "log-genome",
                    "import-comparisons"):
        assert command in result.stdout
    result = runner.invoke(private_cli.app, ["compute-column", "--help"])
    for option in ("--database", "--run-id", "--subject", "--json", "--cache", "--temp", "--log", "--debug"):
        assert option in result.stdout, option
    result = runner.invoke(private_cli.app, ["compute-column", "-d", str(tmp_path / "missing.db"), "-r", "1",
                                             "--subject", "0", "--json", str(tmp_path / "o.json"), "--log", "-"])
    assert result.exit_code == 1 and "does not exist" in str(result.exception)
    result = runner.invoke(private_cli.app, ["prepare-genomes", "-d", str(tmp_path / "missing.db"), "--run-id", "1"])
    assert result.exit_code == 1 and "does not exist" in str(result.exception)
    result = runner.invoke(private_cli.app, ["log-configuration", "-d", str(tmp_path / "new.db"), "--method",
                                             "sourmash", "--program", "panib200", "--version", "0.1.0", "--kmersize",
                                             "31", "--extra", "scaled=300", "--create-db"])
    assert result.exit_code == 0 and (tmp_path / "new.db").is_file()
