"""Public command line interface: ``pyani-plus sourmash`` and ``pyani-plus resume`` on a B200.

Drop-in for the sourmash path of ``pyani_plus/public_cli.py``: ``cli_sourmash`` (:598-639),
``start_and_run_method`` (:115-203), ``run_method`` (:206-329) and ``resume`` (:702-828) keep their
options, log messages, database side effects and error texts.  The reference hands the compute step
to snakemake, which for sourmash schedules ONE job (``column_0``, :232-235) that runs
``.pyani-plus-private-cli compute-column --subject 0``; here that one job is run in-process through
the same ``private_cli.compute_column`` entry point and the same JSON hand-over, so no workflow
engine is needed.  Only the ``local`` executor exists.
"""

from __future__ import annotations

import logging
import sys
import tempfile
from contextlib import nullcontext
from enum import Enum
from pathlib import Path
from typing import Annotated

import typer
from rich.progress import Progress

from pyani_plus_b200 import LOG_FILE, LOG_FILE_DYNAMIC, __version__, db_orm, log_sys_exit, private_cli, setup_logger, tools
from pyani_plus_b200 import run as run_mod
from pyani_plus_b200.db_orm import Session
from pyani_plus_b200.methods import sourmash
from pyani_plus_b200.utils import available_cores, check_db, check_fasta, fasta_file_stats

app = typer.Typer(no_args_is_help=True, context_settings={"help_option_names": ["-h", "--help"]})

# Runs with more genomes than this record their comparisons straight from arrays
# (private_cli.compute_sourmash_bulk) instead of through the per-pair dict + JSON hand-over.
BULK_THRESHOLD = 200


class ToolExecutor(str, Enum):
    """How the compute step is run (the reference also offers slurm through snakemake)."""

    local = "local"
    slurm = "slurm"


REQ_FASTA_DIR = Annotated[Path, typer.Argument(help="Directory of FASTA files (extensions .fas, .fasta, .fna, .fa; "
                                               "optionally gzipped).", show_default=False)]
REQ_DB = Annotated[Path, typer.Option("--database", "-d", help="Path to pyANI-plus SQLite3 database.",
                                      show_default=False, dir_okay=False, file_okay=True)]
OPT_RUN_NAME = Annotated[str | None, typer.Option(help="Run name. Default is 'N genomes using METHOD'.")]
OPT_CREATE_DB = Annotated[bool, typer.Option(help="Create database if does not exist.")]
OPT_EXECUTOR = Annotated[ToolExecutor, typer.Option(help="How should the internal tools be run?")]
OPT_CACHE = Annotated[Path, typer.Option(help="Cache location for sourmash signatures.", file_okay=False)]
OPT_TEMP = Annotated[Path | None, typer.Option(help="Directory to use for intermediate files (kept, for debugging).",
                                               file_okay=False)]
OPT_WTEMP = Annotated[Path | None, typer.Option(help="Directory to use for the JSON hand-over files.",
                                                file_okay=False)]
OPT_LOG = Annotated[Path, typer.Option(help="Where to record log(s). Use '-' for no logging.", dir_okay=False)]
OPT_SCALED = Annotated[int, typer.Option(help="Sets the compression ratio (FracMinHash scaled).", min=1,
                                         rich_help_panel="Method parameters")]
OPT_KMERSIZE = Annotated[int, typer.Option(help="Comparison method k-mer size.", min=1,
                                           rich_help_panel="Method parameters")]
OPT_DEBUG = Annotated[bool, typer.Option(help="Show debugging level logging at the terminal.")]
OPT_RUN_ID = Annotated[int | None, typer.Option("--run-id", "-r", help="Which run from the database "
                                                "(defaults to latest).", show_default=False)]


def version_callback(value: bool) -> None:  # noqa: FBT001
    if value:
        print(f"pyANI-plus (B200 sourmash engine) {__version__}")  # noqa: T201
        raise typer.Exit


@app.callback()
def common(
    version: Annotated[bool, typer.Option("--version", "-v", help="Show tool version (on stdout) and quit.",
                                          callback=version_callback, is_eager=True)] = False,  # noqa: FBT002
) -> None:
    """pyANI-plus ANI analysis: the sourmash method, computed on an NVIDIA B200."""


def _scan_fasta(filename: Path):  # noqa: ANN202
    """Thread-pool worker: ``fasta_file_stats`` with the ValueError returned instead of raised."""
    try:
        return fasta_file_stats(filename)
    except ValueError as err:
        return err


def start_and_run_method(  # noqa: PLR0913, PLR0917
    logger: logging.Logger,
    executor: ToolExecutor,
    cache: Path,
    temp: Path | None,
    workflow_temp: Path | None,
    database: Path,
    log: Path,
    name: str | None,
    method: str,
    fasta: Path,
    tool: tools.ExternalToolData | None,
    *,
    fragsize: int | None = None,
    mode: str | None = None,
    kmersize: int | None = None,
    minmatch: float | None = None,
    extra: str | None = None,
    ctx: run_mod.DistContext | None = None,
) -> int:
    """Record configuration, genomes and a new run in the database, then compute it."""
    fasta_names = check_fasta(logger, fasta)
    with db_orm.connect_to_db(logger, database) as session:
        config = db_orm.db_configuration(
            session, method, "" if tool is None else tool.exe_path.stem, "" if tool is None else tool.version,
            fragsize, mode, kmersize, minmatch, extra, create=True,
        )
        n = len(fasta_names)
        filename_to_md5: dict[Path, str] = {}
        hashes: set[str] = set()
        from concurrent.futures import ThreadPoolExecutor  # noqa: PLC0415

        # md5 + length + description from ONE read per file, files in parallel (all GIL-free C)
        with Progress() as progress, ThreadPoolExecutor(max_workers=max(1, min(16, available_cores()))) as pool:
            scans = pool.map(_scan_fasta, fasta_names)
            for filename in progress.track(fasta_names, description="Indexing FASTAs"):
                scan = next(scans)
                if isinstance(scan, ValueError):
                    log_sys_exit(logger, str(scan))
                md5, stats = scan[0], scan[1:]
                filename_to_md5[filename] = md5
                if md5 in hashes:
                    dups = "\n" + "\n".join(sorted({str(k) for k, v in filename_to_md5.items() if v == md5}))
                    msg = f"Multiple genomes with same MD5 checksum {md5}:{dups}"
                    log_sys_exit(logger, msg)
                hashes.add(md5)
                # one transaction for all genomes: add_run below commits it together with the run
                db_orm.db_genome(logger, session, filename, md5, create=True, stats=stats, commit=False)
        run = db_orm.add_run(
            session, config, cmdline=" ".join(sys.argv), fasta_directory=fasta, status="Initialising",
            name=f"{len(filename_to_md5)} genomes using {method}" if name is None else name, date=None,
            fasta_to_hash=filename_to_md5,
        )
        session.commit()
        msg = f"{method} run setup with {n} genomes in database"
        logger.info(msg)
        return run_method(logger, executor, cache, temp, workflow_temp, filename_to_md5, database, log, session, run,
                          ctx)


def run_method(  # noqa: PLR0913, PLR0917
    logger: logging.Logger,
    executor: ToolExecutor,
    cache: Path,
    temp: Path | None,
    workflow_temp: Path | None,
    filename_to_md5: dict[Path, str],
    database: Path,
    log: Path,  # noqa: ARG001
    session: Session,
    run: db_orm.Run,
    ctx: run_mod.DistContext | None = None,
) -> int:
    """Compute the comparisons the run still lacks and record them in the database.

    How, depending on what is there (the reference hands all of these to snakemake workers,
    ``public_cli.py:206-329``):

    * several GPUs (``torchrun``): every rank takes part (``private_cli.compute_sourmash_distributed``);
    * a fresh small run: the reference's one ``column_0`` job, in-process, through the JSON hand-over;
    * a fresh large run: the same all-vs-all recorded from arrays;
    * a partial run (``resume``): only the queries x subjects block that is missing, as one rectangular call.
    """
    ctx = ctx or run_mod.DistContext()
    run_id = run.run_id
    method = run.configuration.method
    logger.debug("Counting pre-existing comparisons for this run...")
    done = run.comparisons().count()
    n = len(filename_to_md5)
    if done == n**2:
        msg = f"Database already has all {n}²={n**2} {method} comparisons"
        logger.info(msg)
        return 0
    msg = f"Database already has {done} of {n}²={n**2} {method} comparisons, {n**2 - done} needed"
    logger.info(msg)
    if method != "sourmash":
        msg = f"Unknown method {method} for run-id {run_id} in {database}"
        log_sys_exit(logger, msg)
    if executor != ToolExecutor.local:
        msg = f"Executor {executor.value} is not available: the B200 sourmash engine runs in-process (local)"
        log_sys_exit(logger, msg)
    run.status = "Running"
    session.commit()

    def finish(session_: Session, run_: db_orm.Run) -> int:
        have = run_.comparisons().count()
        if have != n**2:
            msg = f"Only have {have} of {n}²={n**2} {method} comparisons needed"  # pragma: no cover
            log_sys_exit(logger, msg)  # pragma: no cover
        if run_.df_identity is None or done:
            run_.cache_comparisons()
        if run_.df_identity is None:
            msg = (f"The {n}x{n} matrices are too large for the database's JSON cache (SQLite stores at most 10^9 "
                   "bytes per text); they will be read from the comparisons table when needed")
            logger.warning(msg)
        run_.status = "Done"
        session_.commit()
        msg = f"Completed {method} run-id {run_id} with {n} genomes in database {database}"
        logger.info(msg)
        return 0

    if ctx.world > 1:
        msg = f"Using {ctx.world} GPUs (one process each)"
        logger.info(msg)
        config = run.configuration
        job = {
            "entries": sorted((md5, str(path)) for path, md5 in filename_to_md5.items()),
            "ksize": int(config.kmersize), "scaled": sourmash.parse_scaled(logger, config.extra),
            "sig_cache": str(cache.absolute() / f"sourmash_k={config.kmersize}_{config.extra}"),
        }
        ctx.broadcast_object(job)
        private_cli.compute_sourmash_distributed(logger, session, job, cache, ctx, run)
        return finish(session, run)

    private_cli.prepare(logger, run, cache)  # builds the .sig cache on the GPU

    if done:  # resume: one rectangular call for the block that is missing
        queries, subjects = private_cli.missing_block(run)
        msg = f"Computing the missing block: {len(queries)} queries x {len(subjects)} subjects"
        logger.info(msg)
        private_cli.compute_sourmash_block(logger, session, run, cache, queries, subjects)
        return finish(session, run)
    if n > BULK_THRESHOLD:
        logger.debug("Recording comparisons from arrays (bulk path)")
        private_cli.compute_sourmash_bulk(logger, session, run, cache)
        return finish(session, run)

    # sourmash: all the columns at once, a single worker -- the reference's `column_0` job
    target = f"{method}.run_{run_id}.column_0.json"
    logger.debug("Using a single worker")
    session.close()  # reduce chance of DB locking
    del run
    with (
        nullcontext(workflow_temp.absolute()) if workflow_temp
        else tempfile.TemporaryDirectory(prefix="pyani-plus_")
    ) as tmp:
        out_path = Path(tmp) / "output"
        out_path.mkdir(parents=True, exist_ok=True)
        json_path = out_path / target
        rc = private_cli.run_compute_column(
            logger, Path(database).absolute(), run_id, "0", json_path,
            cache=cache.absolute(), temp=temp.absolute() if temp else Path("-"),
        )
        if rc:
            msg = f"compute-column returned {rc} for run-id {run_id}"
            log_sys_exit(logger, msg)
        with db_orm.connect_to_db(logger, database) as session2:
            run2 = session2.get_run(run_id)
            if json_path.is_file():
                private_cli.import_json_comparisons(logger, session2, json_path)
            return finish(session2, run2)


def _as_worker(logger: logging.Logger, ctx: run_mod.DistContext, cache: Path) -> int:
    """Ranks other than 0 of a ``torchrun`` launch: no database, no FASTA indexing.  They wait for rank 0 to
    broadcast the job (or "stop" when there is nothing to compute or rank 0 failed early), take their
    share of it, and leave."""
    job = ctx.broadcast_object(None)
    if job == run_mod.JOB_STOP:
        return 0
    private_cli.compute_sourmash_distributed(logger, None, job, cache, ctx)
    return 0


class _Rank0Guard:
    """Makes sure the other ranks are released whatever happens on rank 0 before the job is broadcast."""

    def __init__(self, ctx: run_mod.DistContext) -> None:
        self.ctx, self.sent = ctx, False
        if ctx.world > 1:
            original = ctx.broadcast_object

            def once(obj, src: int = 0):  # noqa: ANN001, ANN202
                self.sent = True
                return original(obj, src)

            ctx.broadcast_object = once  # type: ignore[method-assign]

    def __enter__(self) -> "_Rank0Guard":  # noqa: UP037
        return self

    def __exit__(self, *exc: object) -> None:
        if self.ctx.world > 1 and not self.sent:
            self.ctx.broadcast_object(run_mod.JOB_STOP)
        self.ctx.barrier()
        self.ctx.close()


@app.command("sourmash", rich_help_panel="ANI methods")
def cli_sourmash(  # noqa: PLR0913
    fasta: REQ_FASTA_DIR,
    database: REQ_DB,
    *,
    name: OPT_RUN_NAME = None,
    create_db: OPT_CREATE_DB = False,
    executor: OPT_EXECUTOR = ToolExecutor.local,
    cache: OPT_CACHE = Path(),
    temp: OPT_TEMP = None,
    wtemp: OPT_WTEMP = None,
    log: OPT_LOG = LOG_FILE_DYNAMIC,
    scaled: OPT_SCALED = sourmash.SCALED,  # 1000
    kmersize: OPT_KMERSIZE = sourmash.KMER_SIZE,
    debug: OPT_DEBUG = False,
) -> int:
    """Execute sourmash (FracMinHash max-containment) ANI calculations, logged to a pyANI-plus SQLite3 database."""
    if log == LOG_FILE_DYNAMIC:
        log = Path("-") if executor == ToolExecutor.local else LOG_FILE
    logger = setup_logger(log, terminal_level=logging.DEBUG if debug else logging.INFO)
    ctx = run_mod.DistContext.from_env()  # torchrun: one process per GPU; rank 0 owns the database
    if ctx.rank != 0:
        try:
            return _as_worker(logger, ctx, cache)
        finally:
            ctx.barrier()
            ctx.close()
    with _Rank0Guard(ctx):
        check_db(logger, database, create_db)
        return start_and_run_method(
            logger, executor, cache, temp, wtemp, database, log, name, "sourmash", fasta, tools.get_sourmash(),
            kmersize=kmersize, extra=f"scaled={scaled}", ctx=ctx,
        )


@app.command()
def resume(  # noqa: PLR0913
    database: REQ_DB,
    *,
    run_id: OPT_RUN_ID = None,
    executor: OPT_EXECUTOR = ToolExecutor.local,
    cache: OPT_CACHE = Path(),
    temp: OPT_TEMP = None,
    wtemp: OPT_WTEMP = None,
    log: OPT_LOG = LOG_FILE_DYNAMIC,
    debug: OPT_DEBUG = False,
) -> int:
    """Resume any (partial) run already logged in the database.

    Missing pairwise comparisons are computed and the run is marked as complete; a complete run is
    left alone.  Aborts if the engine version differs from the one recorded for the run.
    """
    if log == LOG_FILE_DYNAMIC:
        log = Path("-") if executor == ToolExecutor.local else LOG_FILE
    logger = setup_logger(log, terminal_level=logging.DEBUG if debug else logging.INFO)
    ctx = run_mod.DistContext.from_env()
    if ctx.rank != 0:
        try:
            return _as_worker(logger, ctx, cache)
        finally:
            ctx.barrier()
            ctx.close()
    with _Rank0Guard(ctx):
        return _resume_rank0(logger, ctx, database, run_id, executor, cache, temp, wtemp, log)


def _resume_rank0(  # noqa: PLR0913, PLR0917
    logger: logging.Logger, ctx: run_mod.DistContext, database: Path, run_id: int | None, executor: ToolExecutor,
    cache: Path, temp: Path | None, wtemp: Path | None, log: Path,
) -> int:
    if database == ":memory:" or not Path(database).is_file():
        msg = f"Database {database} does not exist"
        log_sys_exit(logger, msg)
    with db_orm.connect_to_db(logger, database) as session:
        run = db_orm.load_run(session, run_id)
        if run_id is None:
            run_id = run.run_id
            msg = f"Resuming run-id {run_id}"
            logger.info(msg)
        config = run.configuration
        msg = (
            f"This is a {config.method} run on {run.genomes.count()} genomes, "
            f"using {config.program} version {config.version}"
        )
        logger.info(msg)
        if not run.genomes.count():
            msg = f"No genomes recorded for run-id {run_id}, cannot resume."
            log_sys_exit(logger, msg)
        if config.method != "sourmash":
            msg = f"Unknown method {config.method} for run-id {run_id} in {database}"
            log_sys_exit(logger, msg)
        tool = tools.get_sourmash()
        if tool.exe_path.stem != config.program or tool.version != config.version:
            msg = (
                f"We have {tool.exe_path.stem} version {tool.version}, but"
                f" run-id {run_id} used {config.program} version {config.version} instead."
            )
            log_sys_exit(logger, msg)
        fasta = Path(run.fasta_directory)
        if not fasta.is_dir():
            msg = f"run-id {run_id} used input folder {fasta}, but that is not a directory (now)."
            log_sys_exit(logger, msg)
        filename_to_md5 = {fasta / link.fasta_filename: link.genome_hash for link in run.fasta_hashes}
        for filename, md5 in filename_to_md5.items():
            if not filename.is_file():
                msg = f"run-id {run_id} used {filename} with MD5 {md5} but this FASTA file no longer exists"
                log_sys_exit(logger, msg)
        run.status = "Resuming"
        session.commit()
        return run_method(logger, executor, cache, temp, wtemp, filename_to_md5, database, log, session, run, ctx)


@app.command()
def list_runs(database: REQ_DB) -> int:
    """List the runs defined in a given pyANI-plus SQLite3 database."""
    logger = setup_logger(None)
    if database == ":memory:" or not Path(database).is_file():
        msg = f"Database {database} does not exist"
        log_sys_exit(logger, msg)
    from rich.console import Console  # noqa: PLC0415
    from rich.table import Table  # noqa: PLC0415

    with db_orm.connect_to_db(logger, database) as session:
        runs = session.runs()
        table = Table(title=f"{len(runs)} analysis runs in {database}", row_styles=["dim", ""])
        from rich.text import Text  # noqa: PLC0415

        for col in ("ID", "Date", "Method", "Done", "Null", "Miss", "Total", "Status", "Name"):
            if col == "ID":
                table.add_column(col, justify="right", no_wrap=True)
            elif col in ("Done", "Null", "Miss", "Total"):  # numbers right-aligned under left-aligned headings
                table.add_column(Text(col, justify="left"), justify="right", no_wrap=True)
            else:
                table.add_column(col)
        for run in runs:
            conf = run.configuration
            n = run.genomes.count()
            have = run.comparisons().count()  # counted by SQLite: a large run has 10^8 rows
            null = run.comparisons().null_count()
            table.add_row(
                str(run.run_id), str(run.date.date()), conf.method, str(have - null), str(null),
                str(n**2 - have), f"{n**2}={n}²", run.status, run.name,
            )
    Console().print(table)
    return 0


@app.command()
def export_run(  # noqa: PLR0913
    database: REQ_DB,
    outdir: Annotated[Path, typer.Option(help="Output directory (created if missing).", file_okay=False,
                                         show_default=False)],
    run_id: OPT_RUN_ID = None,
    label: Annotated[str, typer.Option(help="How to label the genomes: md5, filename or stem.")] = "stem",
    log: OPT_LOG = Path("-"),
    *,
    debug: OPT_DEBUG = False,
) -> int:
    """Export any single run: the long-form table ``<method>_run_<run-id>.tsv`` (every comparison with all its
    properties, NA for NULL) and, for a complete run, the matrices ``<method>_<property>.tsv``.

    As the reference's command (public_cli.py:974-1091): a partial run gets its long-form table and then an
    error instead of matrices; an empty run is an error.  Rows are streamed from SQLite in (query, subject) order.
    """
    from math import log as math_log  # noqa: PLC0415

    logger = setup_logger(log, terminal_level=logging.DEBUG if debug else logging.INFO)
    if database == ":memory:" or not Path(database).is_file():
        msg = f"Database {database} does not exist"
        log_sys_exit(logger, msg)
    if not outdir.is_dir():
        msg = f"Output directory {outdir} does not exist, making it."
        logger.warning(msg)
        outdir.mkdir(parents=True)
    with db_orm.connect_to_db(logger, database) as session:
        run = db_orm.load_run(session, run_id, check_empty=True)
        if run_id is None:
            run_id = run.run_id
            msg = f"Exporting run-id {run_id}"
            logger.info(msg)
        method = run.configuration.method
        if label == "md5":
            names = {a.genome_hash: a.genome_hash for a in run.fasta_hashes}
        elif label == "filename":
            names = {a.genome_hash: a.fasta_filename for a in run.fasta_hashes}
        else:
            names = {a.genome_hash: db_orm.filename_stem(a.fasta_filename) for a in run.fasta_hashes}

        def cell(value: float | None) -> str:
            return "NA" if value is None else str(value)

        long_name = f"{method}_run_{run_id}.tsv"
        with (outdir / long_name).open("w") as handle:
            handle.write("#Query\tSubject\tIdentity\tQuery-Cov\tSubject-Cov\tHadamard\ttANI\tAlign-Len\tSim-Errors\n")
            for query, subject, identity, cov_query, cov_subject, aln_length, sim_errors in run.comparisons().values(
                    ("identity", "cov_query", "cov_subject", "aln_length", "sim_errors")):
                hadamard = None if identity is None or cov_query is None else identity * cov_query
                tani = None if hadamard is None else -math_log(hadamard)
                handle.write("\t".join((names[query], names[subject], cell(identity), cell(cov_query), cell(cov_subject),
                                        cell(hadamard), cell(tani), cell(aln_length), cell(sim_errors))) + "\n")
        msg = f"Wrote long-form to {outdir}/{long_name}"
        logger.info(msg)

        run = db_orm.load_run(session, run_id, check_complete=True)  # a partial run stops here
        for matrix, filename in ((run.identities, f"{method}_identity.tsv"),
                                 (run.aln_length, f"{method}_aln_lengths.tsv"),
                                 (run.sim_errors, f"{method}_sim_errors.tsv"),
                                 (run.cov_query, f"{method}_query_cov.tsv"),
                                 (run.hadamard, f"{method}_hadamard.tsv"),
                                 (run.tani, f"{method}_tANI.tsv")):
            try:
                matrix = run.relabelled_matrix(matrix, label)
            except ValueError as err:
                log_sys_exit(logger, str(err))
            matrix.to_csv(outdir / filename, sep="\t")
    msg = f"Wrote matrices to {outdir}/{method}_*.tsv"
    logger.info(msg)
    return 0


if __name__ == "__main__":
    sys.exit(app())  # pragma: no cover
