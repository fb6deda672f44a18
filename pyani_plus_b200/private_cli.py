"""Private (worker) command line interface for the sourmash path.

Drop-in for the parts of ``pyani_plus/private_cli.py`` the path uses; same command names, options,
return codes and error texts:

* ``prepare-genomes`` / ``prepare()``                       reference :677-754
* ``compute-column`` (``--subject 0`` = all columns)         reference :757-973
* ``compute_sourmash()``                                     reference :1803-1902
* ``export_json_db_entries`` / ``import_json_comparisons``  reference :454-614 (JSON schema kept)
* ``log-configuration`` / ``log-genome`` / ``log-run`` / ``import-comparisons``  reference :226-674 (used to set up tests)

Only ``sourmash`` is a known method here: every other method of pyani-plus wraps an external
aligner and is out of scope.
"""

from __future__ import annotations

import json as _json
import logging
import platform
import signal
import sys
import tempfile
from contextlib import nullcontext
from pathlib import Path
from typing import Annotated

import typer

from pyani_plus_b200 import LOG_FILE, db_orm, log_sys_exit, setup_logger, tools
from pyani_plus_b200.db_orm import NoResultFound, Session
from pyani_plus_b200.utils import check_fasta, file_md5sum

app = typer.Typer(context_settings={"help_option_names": ["-h", "--help"]}, no_args_is_help=True)

NO_PATH = Path("-")
RECORDING_FAILED = 2  # return code for successful calculation but failed to save to DB

REQ_DB = Annotated[Path, typer.Option("--database", "-d", help="Path to pyANI-plus SQLite3 database.",
                                      show_default=False, dir_okay=False, file_okay=True)]
OPT_CACHE = Annotated[Path, typer.Option(help="Cache location (sourmash signatures are kept under it).",
                                         file_okay=False, dir_okay=True)]
OPT_DEBUG = Annotated[bool, typer.Option(help="Show debugging level logging at the terminal.")]
OPT_LOG = Annotated[Path, typer.Option(help="Where to record log(s). Use '-' for no logging.",
                                       dir_okay=False, file_okay=True)]
OPT_CREATE_DB = Annotated[bool, typer.Option(help="Create database if does not exist.")]


def _check_tool_version(
    logger: logging.Logger, tool: tools.ExternalToolData, configuration: db_orm.Configuration
) -> None:
    """Confirm the tool and version matches the given configuration (reference :191-223).

    >>> logger = setup_logger(None)
    >>> tool = tools.ExternalToolData(Path("/bin/guestimator"), "1.3")
    >>> config = db_orm.Configuration(method="guessing", program="guestimator", version="1.2")
    >>> _check_tool_version(logger, tool, config)
    Traceback (most recent call last):
    ...
    SystemExit: Run configuration was guestimator 1.2 but we have guestimator 1.3
    """
    if configuration.program != tool.exe_path.stem or configuration.version != tool.version:
        msg = (
            "Run configuration was"
            f" {configuration.program} {configuration.version}"
            f" but we have {tool.exe_path.stem} {tool.version}"
        )
        log_sys_exit(logger, msg)


# ---------------------------------------------------------------------------------------------
# JSON hand-over between the compute step and the database (schema: SURVEY.md 3.4)
# ---------------------------------------------------------------------------------------------
_CONFIG_FIELDS = ("method", "program", "version", "fragsize", "mode", "kmersize", "minmatch", "extra")
_CONFIG_REQUIRED = ("method", "program", "version")
_ROW_REQUIRED = ("query_hash", "subject_hash", "identity")
_ROW_OPTIONAL = ("aln_length", "sim_errors", "cov_query")
_ROW_PRIVATE = ("configuration_id", "uname_system", "uname_release", "uname_machine")  # implied by the envelope


def export_json_db_entries(
    logger: logging.Logger,
    json_filename: Path,
    configuration: db_orm.Configuration,
    db_entries: list[dict[str, str | float | int | None]],
) -> None:
    """Write comparisons as the JSON hand-over file (schema: reference private_cli.py:454-504).

    One envelope per file: the configuration the rows belong to, this machine's uname, and the rows with the
    per-row copies of those two dropped.
    """
    host = platform.uname()
    envelope = {
        "configuration": {field: getattr(configuration, field) for field in _CONFIG_FIELDS},
        "uname": {"system": host.system, "release": host.release, "machine": host.machine},
        "comparisons": [{k: v for k, v in row.items() if k not in _ROW_PRIVATE} for row in db_entries],
    }
    json_filename.write_text(_json.dumps(envelope))
    msg = f"Saved {len(db_entries)} comparisons to {json_filename}"
    logger.debug(msg)


def _json_envelope(logger: logging.Logger, session: Session, json_filename: Path) -> tuple[dict, dict, list] | None:
    """Parse and shape-check a hand-over file: (configuration, uname, rows), or None for an empty file.

    Every malformed input ends the program with the reference's message for that case.
    """
    raw = json_filename.read_bytes()
    if not raw:
        msg = f"JSON file '{json_filename}' is empty"
        logger.debug(msg)
        return None
    try:
        data = _json.loads(raw)
    except ValueError:
        logger.exception("Unable to parse JSON:")
        log_sys_exit(logger, f"JSON file '{json_filename}' invalid")
    shape = {"configuration": dict, "uname": dict, "comparisons": list}
    if not isinstance(data, dict) or any(not isinstance(data.get(k), t) for k, t in shape.items()):
        log_sys_exit(logger, f"JSON file '{json_filename}' does not use the expected structure")
    problems = (
        ("uname incomplete", data["uname"], ("system", "release", "machine")),
        ("configuration incomplete", data["configuration"], _CONFIG_REQUIRED),
    )
    for what, mapping, keys in problems:
        if any(k not in mapping for k in keys):
            session.close()
            log_sys_exit(logger, f"JSON file '{json_filename}' {what}")
    return data["configuration"], data["uname"], data["comparisons"]


def import_json_comparisons(logger: logging.Logger, session: Session, json_filename: Path) -> int:
    """Record the comparisons of a hand-over file in the database (INSERT OR IGNORE); returns how many
    the file held.  The configuration must already be in the database (reference private_cli.py:507-614)."""
    msg = f"Importing {json_filename}"
    logger.debug(msg)
    envelope = _json_envelope(logger, session, json_filename)
    if envelope is None:
        return 0
    configuration, uname, rows = envelope
    try:
        config_id = db_orm.db_configuration(
            session, *(configuration.get(field) for field in _CONFIG_FIELDS), create=False
        ).configuration_id
    except NoResultFound:
        session.close()
        log_sys_exit(logger, f"JSON file '{json_filename}' configuration not in database")
    msg = f"Configuration identifier {config_id} in database"
    logger.debug(msg)
    if not rows:
        session.close()
        msg = f"JSON file '{json_filename}' has no comparisons"
        logger.warning(msg)
        return 0
    if any(key not in row for row in rows for key in _ROW_REQUIRED):
        session.close()
        log_sys_exit(logger, f"JSON file '{json_filename}' comparison(s) incomplete")
    shared = {"configuration_id": config_id, "uname_system": uname["system"], "uname_release": uname["release"],
              "uname_machine": uname["machine"]}
    db_entries = [
        {**shared, **{k: row[k] for k in _ROW_REQUIRED}, **{k: row.get(k) for k in _ROW_OPTIONAL}} for row in rows
    ]
    if not db_orm.insert_comparisons_with_retries(logger, session, db_entries, source=str(json_filename)):
        session.close()  # pragma: no cover
        log_sys_exit(logger, f"Failed to record '{json_filename}' comparisons to database")  # pragma: no cover
    return len(rows)


# ---------------------------------------------------------------------------------------------
# low-level logging commands (database set-up; used by the tests exactly as in the reference)
# ---------------------------------------------------------------------------------------------
def _require_db(database: Path | str, *, create_db: bool) -> None:
    if database != ":memory:" and not create_db and not Path(database).is_file():
        msg = f"Database '{database}' does not exist, but not using --create-db"
        sys.exit(msg)


@app.command(rich_help_panel="Low-level logging")
def log_configuration(  # noqa: PLR0913
    database: REQ_DB,
    method: Annotated[str, typer.Option(help="Method, e.g. sourmash", show_default=False)],
    program: Annotated[str, typer.Option(help="Program, e.g. panib200", show_default=False)],
    version: Annotated[str, typer.Option(help="Program version", show_default=False)],
    *,
    fragsize: Annotated[int | None, typer.Option(help="Fragment size")] = None,
    mode: Annotated[str | None, typer.Option(help="Mode")] = None,
    kmersize: Annotated[int | None, typer.Option(help="k-mer size")] = None,
    minmatch: Annotated[float | None, typer.Option(help="Min-match")] = None,
    extra: Annotated[str | None, typer.Option(help="Method specific setting, e.g. scaled=1000")] = None,
    create_db: OPT_CREATE_DB = False,
) -> int:
    """Log a specific method configuration to database (pre-existing entries are left as is)."""
    logger = setup_logger(None, terminal_level=logging.INFO)
    _require_db(database, create_db=create_db)
    with db_orm.connect_to_db(logger, database) as session:
        config = db_orm.db_configuration(session, method, program, version, fragsize, mode, kmersize, minmatch,
                                         extra, create=True)
        msg = f"Configuration identifier {config.configuration_id}"
        logger.info(msg)
    return 0


@app.command(rich_help_panel="Low-level logging")
def log_genome(
    fasta: Annotated[list[Path], typer.Argument(help="FASTA file(s)", show_default=False)],
    database: REQ_DB,
    *,
    create_db: OPT_CREATE_DB = False,
) -> int:
    """Compute MD5 checksums of given FASTA files, log them to database."""
    logger = setup_logger(None, terminal_level=logging.INFO)
    _require_db(database, create_db=create_db)
    with db_orm.connect_to_db(logger, database) as session:
        for filename in fasta:
            db_orm.db_genome(logger, session, filename, file_md5sum(filename), create=True)
        session.commit()
    msg = f"Processed {len(fasta)} FASTA files"
    logger.info(msg)
    return 0


@app.command(rich_help_panel="Low-level logging")
def log_run(  # noqa: PLR0913, PLR0917
    fasta: Annotated[Path, typer.Argument(help="Directory of FASTA files", show_default=False)],
    database: REQ_DB,
    cmdline: Annotated[str, typer.Option(help="Run command line", show_default=False)],
    status: Annotated[str, typer.Option(help="Run status", show_default=False)],
    name: Annotated[str, typer.Option(help="Run name", show_default=False)],
    method: Annotated[str, typer.Option(help="Method, e.g. sourmash", show_default=False)],
    program: Annotated[str, typer.Option(help="Program", show_default=False)],
    version: Annotated[str, typer.Option(help="Program version", show_default=False)],
    *,
    fragsize: Annotated[int | None, typer.Option(help="Fragment size")] = None,
    mode: Annotated[str | None, typer.Option(help="Mode")] = None,
    kmersize: Annotated[int | None, typer.Option(help="k-mer size")] = None,
    extra: Annotated[str | None, typer.Option(help="Method specific setting")] = None,
    minmatch: Annotated[float | None, typer.Option(help="Min-match")] = None,
    create_db: OPT_CREATE_DB = False,
) -> int:
    """Log a run (and if need be, associated configuration and genome rows)."""
    logger = setup_logger(None, terminal_level=logging.INFO)
    _require_db(database, create_db=create_db)
    msg = f"Logging run to '{database}'"
    logger.info(msg)
    with db_orm.connect_to_db(logger, database) as session:
        config = db_orm.db_configuration(session, method, program, version, fragsize, mode, kmersize, minmatch,
                                         extra, create=True)
        fasta_to_hash = {}
        for filename in check_fasta(logger, fasta):
            md5 = file_md5sum(filename)
            fasta_to_hash[filename] = md5
            db_orm.db_genome(logger, session, filename, md5, create=True)
        run = db_orm.add_run(session, config, cmdline, fasta, status, name, date=None, fasta_to_hash=fasta_to_hash)
        if run.comparisons().count() == len(fasta_to_hash) ** 2:
            run.cache_comparisons()
        run_id = run.run_id
        session.commit()
    msg = f"Run identifier {run_id}"
    logger.info(msg)
    return 0


@app.command(rich_help_panel="Low-level logging")
def import_comparisons(
    database: REQ_DB,
    json: Annotated[list[Path], typer.Argument(help="JSON file(s) of comparisons", show_default=False)],
    *,
    debug: OPT_DEBUG = False,
    log: OPT_LOG = NO_PATH,
) -> int:
    """Import JSON file(s) of pairwise comparisons into the database.

    The database must already hold the matching configuration (log-configuration) and the genomes (log-genome);
    each file carries one configuration, one uname and any number of comparisons (reference
    private_cli.py:618-674: what its workflow calls while a run progresses).
    """
    logger = setup_logger(log, terminal_level=logging.DEBUG if debug else logging.ERROR)
    if database != ":memory:" and not Path(database).is_file():
        msg = f"Database '{database}' does not exist"
        sys.exit(msg)
    msg = f"Logging comparison to '{database}'"
    logger.info(msg)
    with db_orm.connect_to_db(logger, database) as session:
        for table, what in (("configurations", "configurations"), ("genomes", "genomes")):
            if session.execute(f"SELECT COUNT(*) FROM {table}").fetchone()[0] == 0:  # noqa: S608
                msg = f"{database} does not contain any {what}"
                log_sys_exit(logger, msg)
        for filename in json:
            count = import_json_comparisons(logger, session, filename)
            msg = f"Imported {count} from '{filename}'"
            logger.info(msg)
    return 0


# ---------------------------------------------------------------------------------------------
# prepare-genomes
# ---------------------------------------------------------------------------------------------
@app.command()
def prepare_genomes(
    database: REQ_DB,
    run_id: Annotated[int | None, typer.Option(help="Which run to prepare", show_default=False)] = None,
    cache: OPT_CACHE = Path(),
    *,
    debug: OPT_DEBUG = False,
    log: OPT_LOG = NO_PATH,
) -> int:
    """Prepare any intermediate files needed prior to computing ANI values.

    For sourmash this builds the signature files (on the GPU) under the cache directory; use the same
    cache location afterwards with compute-column.
    """
    logger = setup_logger(log, terminal_level=logging.DEBUG if debug else logging.ERROR, plain=True)
    if database != ":memory:" and not Path(database).is_file():
        msg = f"Database '{database}' does not exist"
        log_sys_exit(logger, msg)
    with db_orm.connect_to_db(logger, database) as session:
        run = db_orm.load_run(session, run_id)
        return prepare(logger, run, cache)


def prepare(logger: logging.Logger, run: db_orm.Run, cache: Path) -> int:
    """Call the method's prepare_genomes (if it has one) with a progress bar."""
    n = run.genomes.count()
    done = run.comparisons().count()
    if done == n**2:
        msg = f"Skipping preparation, run already has all {n**2}={n}² pairwise values"
        logger.info(msg)
        return 0
    method = run.configuration.method

    import importlib  # noqa: PLC0415

    try:
        module = importlib.import_module(f"pyani_plus_b200.methods.{method.lower().replace('-', '_')}")
    except ModuleNotFoundError:
        msg = f"Unknown method {method}, check tool version?"
        log_sys_exit(logger, msg)
    if not hasattr(module, "prepare_genomes"):
        msg = f"No per-genome preparation required for {method}"  # pragma: no cover
        logger.info(msg)  # pragma: no cover
        return 0  # pragma: no cover
    msg = f"Preparing {n} genomes under cache '{cache}'"
    logger.info(msg)

    from rich.progress import Progress  # noqa: PLC0415

    with Progress() as progress:
        for _ in progress.track(module.prepare_genomes(logger, run, cache), description="Processing...  ", total=n):
            pass
    logger.debug("Done")
    return 0


# ---------------------------------------------------------------------------------------------
# compute-column
# ---------------------------------------------------------------------------------------------
@app.command()
def compute_column(  # noqa: C901, PLR0912, PLR0913, PLR0915
    database: REQ_DB,
    run_id: Annotated[int, typer.Option("--run-id", "-r", help="Which run from the database", show_default=False)],
    subject: Annotated[str, typer.Option(help="Subject (reference) FASTA filename, MD5 checksum, or index (integer).",
                                         show_default=False)],
    json: Annotated[Path, typer.Option(help="Output JSON filename", show_default=False, dir_okay=False)],
    *,
    cache: OPT_CACHE = Path(),
    temp: Annotated[Path, typer.Option(help="Directory to use for intermediate files ('-' = system temp).",
                                       file_okay=False)] = Path("-"),
    debug: OPT_DEBUG = False,
    log: OPT_LOG = LOG_FILE,
) -> int:
    """Run the method for one column and log pairwise comparisons to JSON for the database.

    Column numbers are one based; 0 means compute all the columns (sourmash only).  sourmash requires
    that prepare-genomes was run first with the same cache location.
    """
    try:
        column = int(subject)
    except ValueError:
        column = -1  # the column specific log file has to wait until the subject is resolved

    def column_logger(col: int) -> logging.Logger:
        col_log = log if log == NO_PATH else Path(str(log)[: -len(log.suffix)] + f".{col}" + log.suffix)
        return setup_logger(col_log, terminal_level=logging.DEBUG if debug else logging.ERROR, plain=True)

    logger = (
        column_logger(column) if column >= 0
        else setup_logger(None, terminal_level=logging.DEBUG if debug else logging.ERROR, plain=True)
    )
    # receive SIGINT (and SLURM's SIGTERM) as KeyboardInterrupt even when non-interactive
    try:
        signal.signal(signal.SIGINT, signal.default_int_handler)
        signal.signal(signal.SIGTERM, signal.default_int_handler)
    except ValueError:  # pragma: no cover  (not the main thread)
        pass
    return run_compute_column(
        logger, database, run_id, subject, json, cache=cache, temp=temp,
        relog=None if column >= 0 else column_logger,
    )


def resolve_subject(  # noqa: PLR0913
    logger: logging.Logger, run_id: int, method: str, hash_to_filename: dict[str, str],
    filename_to_hash: dict[str, str], subject: str,
) -> tuple[int, str]:
    """What ``--subject`` names: (one-based column over the sorted MD5s, MD5), or (0, "") for "all columns".

    Accepted, in this order: an MD5 of the run, a FASTA filename of the run (any directory part ignored),
    a column number (reference private_cli.py:839-866).
    """
    ordered = sorted(hash_to_filename)
    subject_hash = subject if subject in hash_to_filename else filename_to_hash.get(Path(subject).name)
    if subject_hash is not None:
        return ordered.index(subject_hash) + 1, subject_hash
    if not subject.lstrip("+-").isdigit():
        log_sys_exit(
            logger, f"Did not recognise {subject!r} as an MD5 hash, filename, or column number in run-id {run_id}"
        )
    column = int(subject)
    if 1 <= column <= len(ordered):
        return column, ordered[column - 1]
    if column != 0:
        log_sys_exit(
            logger,
            f"Single column should be in range 1 to {len(ordered)},"
            f" or for some methods {0} meaning all columns, but not {subject}",
        )
    if method != "sourmash":
        log_sys_exit(logger, "All columns currently only implemented for sourmash")
    return 0, ""


def run_compute_column(  # noqa: PLR0913
    logger: logging.Logger,
    database: Path,
    run_id: int,
    subject: str,
    json: Path,
    *,
    cache: Path = Path(),
    temp: Path = Path("-"),
    relog=None,  # noqa: ANN001  callable(column) -> logger, used once the column number is known
) -> int:
    """Body of ``compute-column``; also called in-process by ``public_cli.run_method``."""
    msg = f"Starting compute-column for {subject} to {json}"
    logger.debug(msg)
    if database != ":memory:" and not Path(database).is_file():
        log_sys_exit(logger, f"Database '{database}' does not exist")

    with db_orm.connect_to_db(logger, database) as session:
        try:
            run = session.get_run(run_id)
        except NoResultFound:
            log_sys_exit(
                logger, f"Database has no run-id {run_id}. Use the list-runs command for more information."
            )
        method = run.configuration.method
        links = list(run.fasta_hashes)
        filename_to_hash = {link.fasta_filename: link.genome_hash for link in links}
        hash_to_filename = {link.genome_hash: link.fasta_filename for link in links}
        column, subject_hash = resolve_subject(logger, run_id, method, hash_to_filename, filename_to_hash, subject)
        if relog is not None:
            logger = relog(column)
        msg = f"Logging {method} compute-column {column}"
        logger.info(msg)

        # column 0 = everything; a single column only needs the queries it does not have yet
        wanted = set(hash_to_filename)
        if column:
            wanted -= {comp.query_hash for comp in run.comparisons().where_subject(subject_hash)}
        query_hashes = {g.genome_hash: g.length for g in run.genomes if g.genome_hash in wanted}
        if not query_hashes:
            msg = f"No {method} comparisons needed against {subject_hash}"
            logger.info(msg)
            return 0
        if method != "sourmash":
            log_sys_exit(logger, f"Unknown method {method} for run-id {run_id} in {database}")

        fasta_dir = Path(run.fasta_directory)
        if not fasta_dir.is_absolute():
            fasta_dir = (Path(database).parent / fasta_dir).absolute()
        msg = f"FASTA folder {fasta_dir}"
        logger.debug(msg)
        msg = f"Calling {method} for {len(query_hashes)} queries" + ("" if column == 0 else f" vs {subject_hash}.")
        logger.info(msg)

        if temp == Path("-"):
            workdir = tempfile.TemporaryDirectory()
        else:
            kept = temp / f"c{column}"  # one folder per column: no name clashes between workers
            kept.mkdir(exist_ok=True)
            msg = f"Using temp folder {kept}"
            logger.debug(msg)
            workdir = nullcontext(kept)
        with workdir as tmp_dir:
            return compute_sourmash(
                logger, Path(tmp_dir), session, run, json, fasta_dir, hash_to_filename, filename_to_hash,
                query_hashes, subject_hash, cache=cache,
            )


def _signature_cache(logger: logging.Logger, configuration: db_orm.Configuration, cache: Path) -> Path:
    """The run's signature directory under ``cache``; it must exist (prepare-genomes makes it)."""
    sig_cache = cache / f"sourmash_k={configuration.kmersize}_{configuration.extra}"
    if not sig_cache.is_dir():
        log_sys_exit(
            logger, f"Missing sourmash signatures directory '{sig_cache}' - check cache setting '{cache}'."
        )
    return sig_cache


def compute_sourmash(  # noqa: PLR0913, PLR0917
    logger: logging.Logger,
    tmp_dir: Path,
    session: Session,
    run: db_orm.Run,
    json_filename: Path,
    fasta_dir: Path,  # noqa: ARG001
    hash_to_filename: dict[str, str],  # noqa: ARG001
    filename_to_hash: dict[str, str],  # noqa: ARG001
    query_hashes: dict[str, int],
    subject_hash: str,
    *,
    cache: Path = Path(),
) -> int:
    """Queries vs one subject -- or all-vs-all when ``subject_hash == ""`` -- for sourmash, logged to JSON.

    identity := max-containment ANI, cov_query := query-containment ANI (reference private_cli.py:1875-1887).
    The GPU answers the whole block at once, so the rows are collected in one pass and the file is written
    once (the reference re-dumps it after every 100,000 rows of a slow subprocess).
    """
    configuration = run.configuration
    tool = tools.get_sourmash()
    _check_tool_version(logger, tool, configuration)
    sig_cache = _signature_cache(logger, configuration, cache)

    from pyani_plus_b200.methods import sourmash  # noqa: PLC0415

    host = platform.uname()
    shared = {"configuration_id": configuration.configuration_id, "uname_system": host.system,
              "uname_release": host.release, "uname_machine": host.machine}
    subjects = {subject_hash} if subject_hash else set(query_hashes)
    db_entries: list[dict[str, str | float | int | None]] = []
    try:
        db_entries.extend(
            {"query_hash": q, "subject_hash": s, "identity": max_containment, "cov_query": q_containment, **shared}
            for q, s, q_containment, max_containment in sourmash.compute_sourmash_tile(
                logger, tool, subjects, set(query_hashes), sig_cache, tmp_dir)
        )
    except KeyboardInterrupt:  # pragma: no cover
        msg = f"Interrupted with {len(db_entries)} completed sourmash comparisons"
        logger.error(msg)  # noqa: TRY400
        run.status = "Worker interrupted"
        session.commit()
    try:
        export_json_db_entries(logger, json_filename, configuration, db_entries)
    except Exception:  # pragma: no cover  # noqa: BLE001
        logger.exception("Unexpected exception saving JSON:")
        return RECORDING_FAILED
    return 0


def compute_sourmash_block(  # noqa: PLR0913
    logger: logging.Logger, session: Session, run: db_orm.Run, cache: Path, queries: set[str], subjects: set[str],
) -> tuple[list[str], list[str], object, object]:
    """One queries x subjects block recorded straight from arrays (no per-pair dict, no JSON).

    Same row semantics as ``compute_sourmash`` + ``import_json_comparisons`` (ordered rows, identity :=
    max-containment ANI, cov_query := query-containment ANI, NULL where there is no common hash, INSERT OR
    IGNORE).  ``queries == subjects`` is the all-vs-all call (inverted-index or probing K2, chosen from the
    data); anything else is a rectangular call served by the probing kernel -- what ``resume`` uses to
    compute only what a partial run lacks (reference private_cli.py:879-898).
    Returns (sorted queries, sorted subjects, identity, cov_query).
    """
    configuration = run.configuration
    _check_tool_version(logger, tools.get_sourmash(), configuration)
    sig_cache = _signature_cache(logger, configuration, cache)

    from pyani_plus_b200.methods import sourmash  # noqa: PLC0415

    q_sorted, s_sorted, _, identity, cov_query = sourmash.tile_arrays(logger, subjects, queries, sig_cache, None)
    if not db_orm.insert_comparison_arrays(logger, session, configuration.configuration_id, q_sorted, s_sorted,
                                           identity, cov_query):
        log_sys_exit(logger, "Failed to record comparisons to database")  # pragma: no cover
    return q_sorted, s_sorted, identity, cov_query


def compute_sourmash_bulk(logger: logging.Logger, session: Session, run: db_orm.Run, cache: Path) -> int:
    """All-vs-all for the whole run through ``compute_sourmash_block``; returns the
    number of ordered pairs computed."""
    hashes = {link.genome_hash for link in run.fasta_hashes}
    fresh = run.comparisons().count() == 0
    q_sorted, s_sorted, identity, cov_query = compute_sourmash_block(logger, session, run, cache, hashes, hashes)
    if fresh:  # the matrices just recorded ARE the run's matrices: no need to read N^2 rows back
        run.cache_comparisons(computed=(q_sorted, identity, cov_query))
    return len(q_sorted) * len(s_sorted)


def missing_block(run: db_orm.Run) -> tuple[set[str], set[str]]:
    """The smallest queries x subjects block that covers every comparison the run still lacks.

    Counted inside SQLite (one GROUP BY over the run's comparisons), then only the subjects that are short of a
    full column are looked at row by row: resuming a run of 10,000 genomes does not build 10^8 Python objects.
    """
    hashes = sorted(link.genome_hash for link in run.fasta_hashes)
    everyone = set(hashes)
    session = run._session  # noqa: SLF001
    params = (run.configuration_id, run.run_id, run.run_id)
    have = dict(session.execute("SELECT comparisons.subject_hash, COUNT(*)" + db_orm._RUN_JOIN  # noqa: SLF001
                                + " GROUP BY comparisons.subject_hash", params).fetchall())
    subjects = {h for h in hashes if have.get(h, 0) < len(hashes)}
    queries: set[str] = set()
    for subject in subjects:
        if len(queries) == len(everyone):
            break
        present = {row[0] for row in session.execute(
            "SELECT comparisons.query_hash" + db_orm._RUN_JOIN + " AND comparisons.subject_hash = ?",  # noqa: SLF001
            (*params, subject))} if have.get(subject, 0) else set()
        queries |= everyone - present
    return queries, subjects


def compute_sourmash_distributed(  # noqa: PLR0913
    logger: logging.Logger, session: Session | None, run_job: dict, cache: Path, ctx, run: db_orm.Run | None = None,  # noqa: ANN001
) -> int:
    """All-vs-all over every rank of a ``torchrun`` launch (``run.all_vs_all_files``): each rank sketches a
    slice of the genomes (honouring the shared ``.sig`` cache), the sketches are exchanged once, the pair work
    is sharded, and rank 0 -- the only rank with a database session -- records the N^2 rows.

    ``run_job`` is the picklable description rank 0 broadcasts: entries [(md5, FASTA path)], ksize, scaled,
    sig cache directory.  Returns the number of ordered pairs (0 on the other ranks).
    """
    from pyani_plus_b200 import run as run_mod  # noqa: PLC0415

    sig_cache = Path(run_job["sig_cache"])
    if ctx.rank == 0:
        sig_cache.mkdir(parents=True, exist_ok=True)
    ctx.barrier()
    result = run_mod.all_vs_all_files(logger, ctx, [tuple(e) for e in run_job["entries"]], run_job["ksize"],
                                      run_job["scaled"], sig_cache)
    if ctx.rank != 0 or result is None:
        return 0
    hashes, _, _, identity, cov_query = result
    assert session is not None and run is not None  # noqa: S101
    fresh = run.comparisons().count() == 0
    if not db_orm.insert_comparison_arrays(logger, session, run.configuration.configuration_id, hashes, hashes,
                                           identity, cov_query):
        log_sys_exit(logger, "Failed to record comparisons to database")  # pragma: no cover
    if fresh:
        run.cache_comparisons(computed=(hashes, identity, cov_query))
    return len(hashes) ** 2


if __name__ == "__main__":
    sys.exit(app())  # pragma: no cover
