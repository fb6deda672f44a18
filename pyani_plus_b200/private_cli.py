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
from itertools import batched
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
def export_json_db_entries(
    logger: logging.Logger,
    json_filename: Path,
    configuration: db_orm.Configuration,
    db_entries: list[dict[str, str | float | int | None]],
) -> None:
    """Serialise DB entries for recording in JSON for later import.

    The entries must all belong to the given configuration and to this machine (uname).
    """
    uname = platform.uname()
    unwanted = {"configuration_id", "uname_system", "uname_release", "uname_machine"}
    payload = {
        "configuration": {
            "method": configuration.method,
            "program": configuration.program,
            "version": configuration.version,
            "fragsize": configuration.fragsize,
            "mode": configuration.mode,
            "kmersize": configuration.kmersize,
            "minmatch": configuration.minmatch,
            "extra": configuration.extra,
        },
        "uname": {"system": uname.system, "release": uname.release, "machine": uname.machine},
        "comparisons": [{k: v for k, v in e.items() if k not in unwanted} for e in db_entries],
    }
    with json_filename.open("w") as handle:
        handle.write(_json.dumps(payload))
    msg = f"Saved {len(db_entries)} comparisons to {json_filename}"
    logger.debug(msg)


def import_json_comparisons(logger: logging.Logger, session: Session, json_filename: Path) -> int:  # noqa: PLR0915
    """Import a JSON file of comparisons into the database (INSERT OR IGNORE)."""
    msg = f"Importing {json_filename}"
    logger.debug(msg)
    with json_filename.open("rb") as handle:
        raw = handle.read()
    if not raw:
        msg = f"JSON file '{json_filename}' is empty"
        logger.debug(msg)
        return 0
    try:
        data = _json.loads(raw)
    except ValueError:
        logger.exception("Unable to parse JSON:")
        msg = f"JSON file '{json_filename}' invalid"
        log_sys_exit(logger, msg)
    del raw
    if (
        not isinstance(data, dict)
        or not isinstance(data.get("configuration"), dict)
        or not isinstance(data.get("uname"), dict)
        or not isinstance(data.get("comparisons"), list)
    ):
        msg = f"JSON file '{json_filename}' does not use the expected structure"
        log_sys_exit(logger, msg)
    configuration = data["configuration"]
    comparisons = data["comparisons"]
    try:
        uname_system = data["uname"]["system"]
        uname_release = data["uname"]["release"]
        uname_machine = data["uname"]["machine"]
    except KeyError:
        msg = f"JSON file '{json_filename}' uname incomplete"
        session.close()
        log_sys_exit(logger, msg)
    del data
    try:
        config_id = db_orm.db_configuration(
            session=session,
            method=configuration["method"],
            program=configuration["program"],
            version=configuration["version"],
            fragsize=configuration.get("fragsize", None),
            mode=configuration.get("mode", None),
            kmersize=configuration.get("kmersize", None),
            minmatch=configuration.get("minmatch", None),
            extra=configuration.get("extra", None),
            create=False,
        ).configuration_id
    except KeyError:
        msg = f"JSON file '{json_filename}' configuration incomplete"
        session.close()
        log_sys_exit(logger, msg)
    except NoResultFound:
        msg = f"JSON file '{json_filename}' configuration not in database"
        session.close()
        log_sys_exit(logger, msg)
    msg = f"Configuration identifier {config_id} in database"
    logger.debug(msg)
    if not comparisons:
        msg = f"JSON file '{json_filename}' has no comparisons"
        session.close()
        logger.warning(msg)
        return 0
    try:
        db_entries = [
            {
                "query_hash": row["query_hash"],
                "subject_hash": row["subject_hash"],
                "identity": row["identity"],
                "aln_length": row.get("aln_length", None),
                "sim_errors": row.get("sim_errors", None),
                "cov_query": row.get("cov_query", None),
                "configuration_id": config_id,
                "uname_system": uname_system,
                "uname_release": uname_release,
                "uname_machine": uname_machine,
            }
            for row in comparisons
        ]
    except KeyError:
        msg = f"JSON file '{json_filename}' comparison(s) incomplete"
        session.close()
        log_sys_exit(logger, msg)
    if not db_orm.insert_comparisons_with_retries(logger, session, db_entries, source=str(json_filename)):
        msg = f"Failed to record '{json_filename}' comparisons to database"  # pragma: no cover
        session.close()  # pragma: no cover
        log_sys_exit(logger, msg)  # pragma: no cover
    return len(comparisons)


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
) -> int:
    """Import JSON file(s) of pairwise comparisons into the database."""
    logger = setup_logger(None, terminal_level=logging.DEBUG if debug else logging.INFO)
    if database != ":memory:" and not Path(database).is_file():
        msg = f"Database '{database}' does not exist"
        log_sys_exit(logger, msg)
    total = 0
    with db_orm.connect_to_db(logger, database) as session:
        for filename in json:
            total += import_json_comparisons(logger, session, filename)
    msg = f"Imported {total} from {len(json)} JSON files"
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


def run_compute_column(  # noqa: C901, PLR0912, PLR0913, PLR0915
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
        msg = f"Database '{database}' does not exist"
        log_sys_exit(logger, msg)

    with db_orm.connect_to_db(logger, database) as session:
        try:
            run = session.get_run(run_id)
        except NoResultFound:
            msg = f"Database has no run-id {run_id}. Use the list-runs command for more information."
            log_sys_exit(logger, msg)
        config = run.configuration
        method = config.method
        filename_to_hash = {_.fasta_filename: _.genome_hash for _ in run.fasta_hashes}
        hash_to_filename = {_.genome_hash: _.fasta_filename for _ in run.fasta_hashes}
        n = len(hash_to_filename)

        if subject in hash_to_filename:
            subject_hash = subject
            column = sorted(hash_to_filename).index(subject_hash) + 1
        elif Path(subject).name in filename_to_hash:
            subject_hash = filename_to_hash[Path(subject).name]
            column = sorted(hash_to_filename).index(subject_hash) + 1
        else:
            try:
                column = int(subject)
            except ValueError:
                msg = f"Did not recognise {subject!r} as an MD5 hash, filename, or column number in run-id {run_id}"
                log_sys_exit(logger, msg)
            if 0 < column <= n:
                subject_hash = sorted(hash_to_filename)[column - 1]
            elif column == 0:
                if method == "sourmash":
                    subject_hash = ""
                else:
                    msg = "All columns currently only implemented for sourmash"
                    log_sys_exit(logger, msg)
            else:
                msg = (
                    f"Single column should be in range 1 to {n},"
                    f" or for some methods {0} meaning all columns, but not {subject}"
                )
                log_sys_exit(logger, msg)

        if relog is not None:
            logger = relog(column)
        msg = f"Logging {method} compute-column {column}"
        logger.info(msg)

        if column == 0:
            query_hashes = {_.genome_hash: _.length for _ in run.genomes}  # assume all needed
        else:
            missing = set(hash_to_filename).difference(
                comp.query_hash for comp in run.comparisons().where_subject(subject_hash)
            )
            query_hashes = {_.genome_hash: _.length for _ in run.genomes if _.genome_hash in missing}
        if not query_hashes:
            msg = f"No {method} comparisons needed against {subject_hash}"
            logger.info(msg)
            return 0

        try:
            compute = {"sourmash": compute_sourmash}[method]
        except KeyError:
            msg = f"Unknown method {method} for run-id {run_id} in {database}"
            log_sys_exit(logger, msg)

        fasta_dir = Path(run.fasta_directory)
        if not fasta_dir.is_absolute():
            fasta_dir = (Path(database).parent / fasta_dir).absolute()
        msg = f"FASTA folder {fasta_dir}"
        logger.debug(msg)

        tmp: Path | None = None if temp == Path("-") else temp
        if tmp:
            tmp = tmp / f"c{column}"  # avoid worries about name clashes
            tmp.mkdir(exist_ok=True)
            msg = f"Using temp folder {tmp}"
            logger.debug(msg)
        msg = (
            f"Calling {method} for {len(query_hashes)} queries"
            if column == 0
            else f"Calling {method} for {len(query_hashes)} queries vs {subject_hash}."
        )
        logger.info(msg)

        with nullcontext(tmp) if tmp else tempfile.TemporaryDirectory() as tmp_dir:
            return compute(
                logger, Path(tmp_dir), session, run, json, fasta_dir, hash_to_filename, filename_to_hash,
                query_hashes, subject_hash, cache=cache,
            )


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
    """Run many-vs-subject (or all-vs-all when ``subject_hash == ""``) for sourmash and log to JSON.

    Maps identity := max-containment ANI, cov_query := query-containment ANI (reference :1875-1887).
    """
    uname = platform.uname()
    configuration = run.configuration
    tool = tools.get_sourmash()
    _check_tool_version(logger, tool, configuration)
    config_id = configuration.configuration_id

    from pyani_plus_b200.methods import sourmash  # noqa: PLC0415

    sig_cache = cache / f"sourmash_k={configuration.kmersize}_{configuration.extra}"
    if not sig_cache.is_dir():
        msg = f"Missing sourmash signatures directory '{sig_cache}' - check cache setting '{cache}'."
        log_sys_exit(logger, msg)

    db_entries: list[dict[str, str | float | int | None]] = []
    try:
        for batch in batched(
            sourmash.compute_sourmash_tile(
                logger, tool, {subject_hash} if subject_hash else set(query_hashes), set(query_hashes),
                sig_cache, tmp_dir,
            ),
            100000,
        ):
            logger.debug("Computed batch, about to log to database.")
            db_entries.extend(
                {
                    "query_hash": q,
                    "subject_hash": s,
                    "identity": max_containment,
                    "cov_query": q_containment,
                    "configuration_id": config_id,
                    "uname_system": uname.system,
                    "uname_release": uname.release,
                    "uname_machine": uname.machine,
                }
                for q, s, q_containment, max_containment in batch
            )
    except KeyboardInterrupt:  # pragma: no cover
        msg = f"Interrupted with {len(db_entries)} completed sourmash comparisons"
        logger.error(msg)  # noqa: TRY400
        run.status = "Worker interrupted"
        session.commit()
    # (the reference re-dumps the whole JSON after every 100k rows; the GPU finishes all N^2 pairs
    # before the first row is yielded, so one dump at the end records the same file)
    try:
        export_json_db_entries(logger, json_filename, configuration, db_entries)
    except Exception:  # pragma: no cover  # noqa: BLE001
        logger.exception("Unexpected exception saving JSON:")
        return RECORDING_FAILED
    return 0


def compute_sourmash_bulk(logger: logging.Logger, session: Session, run: db_orm.Run, cache: Path) -> int:
    """All-vs-all for the whole run, recorded straight from arrays (no per-pair dict, no JSON).

    Same row semantics as ``compute_sourmash`` + ``import_json_comparisons`` (N^2 ordered rows,
    identity := max-containment ANI, cov_query := query-containment ANI, NULL where there is no common
    hash, INSERT OR IGNORE), for runs too large for the dict / JSON hand-over.
    Returns the number of ordered pairs computed.
    """
    configuration = run.configuration
    tool = tools.get_sourmash()
    _check_tool_version(logger, tool, configuration)

    from pyani_plus_b200.methods import sourmash  # noqa: PLC0415

    sig_cache = cache / f"sourmash_k={configuration.kmersize}_{configuration.extra}"
    if not sig_cache.is_dir():
        msg = f"Missing sourmash signatures directory '{sig_cache}' - check cache setting '{cache}'."
        log_sys_exit(logger, msg)
    hashes = {_.genome_hash for _ in run.fasta_hashes}
    queries, subjects, _, identity, cov_query = sourmash.tile_arrays(logger, hashes, hashes, sig_cache, None)
    if not db_orm.insert_comparison_arrays(logger, session, configuration.configuration_id, queries, subjects,
                                           identity, cov_query):
        msg = "Failed to record comparisons to database"  # pragma: no cover
        log_sys_exit(logger, msg)  # pragma: no cover
    return len(queries) * len(subjects)


if __name__ == "__main__":
    sys.exit(app())  # pragma: no cover
