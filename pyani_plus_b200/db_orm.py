"""Run / Comparison records of the sourmash path on stdlib ``sqlite3``.

Mirrors ``pyani_plus/db_orm.py`` for everything the path touches: the five tables with the schema
SQLAlchemy generates there (``db_orm.py:67-348``; naming convention :77-85), the record classes
``Genome``, ``Configuration``, ``Comparison``, ``Run``, ``RunGenomeAssociation`` with the same attribute
names, and the helpers ``connect_to_db`` (:635-702), ``db_configuration`` (:705-768), ``db_genome``
(:771-877), ``add_run`` (:880-915), ``load_run`` (:918-974), ``db_comparison`` (:977-1041) and
``insert_comparisons_with_retries`` (:1044-1114, ``INSERT OR IGNORE``).  SQLAlchemy is not available
in this image, so the (small) object layer is written directly on ``sqlite3``; a database written
here is meant to be readable by the reference and vice versa.

Differences by design: ``Run.cache_comparisons`` uses a dict for the md5 -> row lookup (the
reference's ``list.index`` is O(N) per comparison, O(N^3) overall, ``db_orm.py:434-436``), and there
is an array-backed bulk insert for large runs.
"""

from __future__ import annotations

import datetime
import json
import logging
import platform
import sqlite3
from io import StringIO
from math import log, nan
from pathlib import Path
from time import sleep
from typing import TYPE_CHECKING, Any

from pyani_plus_b200 import log_sys_exit
from pyani_plus_b200.utils import filename_stem

if TYPE_CHECKING:
    from pandas import DataFrame

SCHEMA = """
CREATE TABLE IF NOT EXISTS genomes (
	genome_hash VARCHAR NOT NULL,
	path VARCHAR NOT NULL,
	length INTEGER NOT NULL,
	description VARCHAR NOT NULL,
	CONSTRAINT pk_genomes PRIMARY KEY (genome_hash)
);
CREATE TABLE IF NOT EXISTS configurations (
	configuration_id INTEGER NOT NULL,
	method VARCHAR NOT NULL,
	program VARCHAR NOT NULL,
	version VARCHAR NOT NULL,
	fragsize INTEGER,
	mode VARCHAR,
	kmersize INTEGER,
	minmatch FLOAT,
	extra VARCHAR,
	CONSTRAINT pk_configurations PRIMARY KEY (configuration_id),
	CONSTRAINT uq_configurations_method UNIQUE (method, program, version, fragsize, mode, kmersize, minmatch, extra)
);
CREATE TABLE IF NOT EXISTS runs (
	run_id INTEGER NOT NULL,
	configuration_id INTEGER NOT NULL,
	cmdline VARCHAR NOT NULL,
	fasta_directory VARCHAR NOT NULL,
	date DATETIME NOT NULL,
	status VARCHAR NOT NULL,
	name VARCHAR NOT NULL,
	df_identity VARCHAR,
	df_cov_query VARCHAR,
	df_aln_length VARCHAR,
	df_sim_errors VARCHAR,
	df_hadamard VARCHAR,
	CONSTRAINT pk_runs PRIMARY KEY (run_id),
	CONSTRAINT fk_runs_configuration_id_configurations FOREIGN KEY(configuration_id) REFERENCES configurations (configuration_id)
);
CREATE TABLE IF NOT EXISTS comparisons (
	comparison_id INTEGER NOT NULL,
	query_hash VARCHAR NOT NULL,
	subject_hash VARCHAR NOT NULL,
	configuration_id INTEGER NOT NULL,
	identity FLOAT,
	aln_length INTEGER,
	sim_errors INTEGER,
	cov_query FLOAT,
	cov_subject FLOAT,
	uname_system VARCHAR NOT NULL,
	uname_release VARCHAR NOT NULL,
	uname_machine VARCHAR NOT NULL,
	CONSTRAINT pk_comparisons PRIMARY KEY (comparison_id),
	CONSTRAINT uq_comparisons_query_hash UNIQUE (query_hash, subject_hash, configuration_id),
	CONSTRAINT fk_comparisons_query_hash_genomes FOREIGN KEY(query_hash) REFERENCES genomes (genome_hash),
	CONSTRAINT fk_comparisons_subject_hash_genomes FOREIGN KEY(subject_hash) REFERENCES genomes (genome_hash),
	CONSTRAINT fk_comparisons_configuration_id_configurations FOREIGN KEY(configuration_id) REFERENCES configurations (configuration_id)
);
CREATE TABLE IF NOT EXISTS runs_genomes (
	genome_hash VARCHAR NOT NULL,
	run_id INTEGER NOT NULL,
	fasta_filename VARCHAR NOT NULL,
	CONSTRAINT pk_runs_genomes PRIMARY KEY (genome_hash, run_id),
	CONSTRAINT fk_runs_genomes_genome_hash_genomes FOREIGN KEY(genome_hash) REFERENCES genomes (genome_hash),
	CONSTRAINT fk_runs_genomes_run_id_runs FOREIGN KEY(run_id) REFERENCES runs (run_id)
);
"""

# The reference caches each N x N matrix of a run as ONE JSON text in the runs table (db_orm.py:393-466); SQLite's
# SQLITE_MAX_LENGTH is 10^9 bytes, a float cell is ~20 bytes of JSON: runs above ~7,000 genomes are not cached.
MATRIX_CACHE_MAX_BYTES = 900_000_000
MATRIX_JSON_BYTES_PER_CELL = 20

COMPARISON_COLUMNS = (
    "query_hash", "subject_hash", "configuration_id", "identity", "aln_length", "sim_errors",
    "cov_query", "cov_subject", "uname_system", "uname_release", "uname_machine",
)


class NoResultFound(Exception):  # noqa: N818
    """Raised where the reference raises ``sqlalchemy.exc.NoResultFound``."""


class QueryList(list):
    """A list with ``.count()`` (the reference's dynamic relationships are counted that way)."""

    def count(self, *args: Any) -> int:  # type: ignore[override]
        return len(self) if not args else super().count(*args)

    def one(self) -> Any:
        if len(self) != 1:
            msg = f"Expected exactly one row, found {len(self)}"
            raise NoResultFound(msg)
        return self[0]

    def first(self) -> Any:
        return self[0] if self else None


class Session:
    """Thin session over a sqlite3 connection (commit flushes attribute changes of loaded runs)."""

    def __init__(self, conn: sqlite3.Connection, dbpath: Path | str) -> None:
        self.conn = conn
        self.dbpath = dbpath
        self._runs: dict[int, Run] = {}

    def __enter__(self) -> "Session":  # noqa: UP037
        return self

    def __exit__(self, *exc: object) -> None:
        self.close()

    def execute(self, sql: str, params: tuple | dict = ()) -> sqlite3.Cursor:
        return self.conn.execute(sql, params)

    def executemany(self, sql: str, rows: Any) -> sqlite3.Cursor:
        return self.conn.executemany(sql, rows)

    def commit(self) -> None:
        for run in self._runs.values():
            run._flush()  # noqa: SLF001
        self.conn.commit()

    def close(self) -> None:
        try:
            self.conn.close()
        except sqlite3.Error:  # pragma: no cover
            pass

    # ---- the handful of queries the path needs -------------------------------------------
    def runs(self) -> QueryList:
        rows = self.execute("SELECT run_id FROM runs ORDER BY run_id").fetchall()
        return QueryList(self.get_run(r[0]) for r in rows)

    def get_run(self, run_id: int) -> "Run":  # noqa: UP037
        if run_id in self._runs:
            return self._runs[run_id]
        row = self.execute(
            "SELECT run_id, configuration_id, cmdline, fasta_directory, date, status, name, df_identity,"
            " df_cov_query, df_aln_length, df_sim_errors, df_hadamard FROM runs WHERE run_id = ?", (run_id,)
        ).fetchone()
        if row is None:
            msg = f"No run with run_id {run_id}"
            raise NoResultFound(msg)
        run = Run(*row)
        run._attach(self)  # noqa: SLF001
        return run

    def get_configuration(self, configuration_id: int) -> "Configuration":  # noqa: UP037
        row = self.execute(
            "SELECT configuration_id, method, program, version, fragsize, mode, kmersize, minmatch, extra"
            " FROM configurations WHERE configuration_id = ?", (configuration_id,)
        ).fetchone()
        if row is None:
            msg = f"No configuration with configuration_id {configuration_id}"
            raise NoResultFound(msg)
        return Configuration(*row)

    def get_genome(self, genome_hash: str) -> "Genome | None":  # noqa: UP037
        row = self.execute(
            "SELECT genome_hash, path, length, description FROM genomes WHERE genome_hash = ?", (genome_hash,)
        ).fetchone()
        return Genome(*row) if row else None


class Genome:
    """An input genome, identified by the MD5 of its (decompressed) FASTA file."""

    def __init__(self, genome_hash: str, path: str, length: int, description: str) -> None:
        self.genome_hash = genome_hash
        self.path = path
        self.length = length
        self.description = description

    def __repr__(self) -> str:
        return (
            f"Genome(genome_hash={self.genome_hash!r}, path={self.path!r},"
            f" length={self.length}, description={self.description!r})"
        )


class Configuration:
    """Method, tool, version and parameters shared by the comparisons of a run."""

    def __init__(  # noqa: PLR0913
        self, configuration_id: int | None = None, method: str = "", program: str = "", version: str = "",
        fragsize: int | None = None, mode: str | None = None, kmersize: int | None = None,
        minmatch: float | None = None, extra: str | None = None,
    ) -> None:
        self.configuration_id = configuration_id
        self.method = method
        self.program = program
        self.version = version
        self.fragsize = fragsize
        self.mode = mode
        self.kmersize = kmersize
        self.minmatch = minmatch
        self.extra = extra

    def __repr__(self) -> str:
        return (
            f"Configuration(configuration_id={self.configuration_id},"
            f" program={self.program!r}, version={self.version!r},"
            f" fragsize={self.fragsize}, mode={self.mode!r},"
            f" kmersize={self.kmersize}, minmatch={self.minmatch},"
            f" extra={self.extra!r})"
        )


class RunGenomeAssociation:
    """Link between a run and one of its genomes (file name within the run's FASTA directory)."""

    def __init__(self, genome_hash: str, run_id: int, fasta_filename: str) -> None:
        self.genome_hash = genome_hash
        self.run_id = run_id
        self.fasta_filename = fasta_filename


class Comparison:
    """One ordered (query, subject) comparison under a configuration."""

    def __init__(  # noqa: PLR0913
        self, comparison_id: int | None, query_hash: str, subject_hash: str, configuration_id: int,
        identity: float | None, aln_length: int | None, sim_errors: int | None, cov_query: float | None,
        cov_subject: float | None, uname_system: str, uname_release: str, uname_machine: str,
    ) -> None:
        self.comparison_id = comparison_id
        self.query_hash = query_hash
        self.subject_hash = subject_hash
        self.configuration_id = configuration_id
        self.identity = identity
        self.aln_length = aln_length
        self.sim_errors = sim_errors
        self.cov_query = cov_query
        self.cov_subject = cov_subject
        self.uname_system = uname_system
        self.uname_release = uname_release
        self.uname_machine = uname_machine

    def __repr__(self) -> str:
        return (
            f"Comparison(comparison_id={self.comparison_id!r}, "
            f"query_hash={self.query_hash!r}, "
            f"subject_hash={self.subject_hash!r}, "
            f"configuration_id={self.configuration_id!r}, "
            f"identity={self.identity}, "
            f"aln_length={self.aln_length}, "
            f"sim_errors={self.sim_errors}, "
            f"cov_query={self.cov_query}, "
            f"cov_subject={self.cov_subject}, "
            f"uname_system={self.uname_system!r}, "
            f"uname_release={self.uname_release!r}, "
            f"uname_machine={self.uname_machine!r})"
        )


_RUN_JOIN = (
    " FROM comparisons"
    " JOIN runs_genomes AS run_query ON comparisons.query_hash = run_query.genome_hash"
    " JOIN runs_genomes AS run_subject ON comparisons.subject_hash = run_subject.genome_hash"
    " WHERE comparisons.configuration_id = ? AND run_query.run_id = ? AND run_subject.run_id = ?"
)


class RunComparisons:
    """The comparisons of a run (lazy: ``count()`` is one SQL query, iteration loads the rows)."""

    def __init__(self, run: "Run", subject_hash: str | None = None) -> None:  # noqa: UP037
        self.run = run
        self.subject_hash = subject_hash

    def _where(self) -> tuple[str, tuple]:
        sql = _RUN_JOIN
        params: tuple = (self.run.configuration_id, self.run.run_id, self.run.run_id)
        if self.subject_hash is not None:
            sql += " AND comparisons.subject_hash = ?"
            params += (self.subject_hash,)
        return sql, params

    def count(self) -> int:
        sql, params = self._where()
        return int(self.run._session.execute("SELECT COUNT(*)" + sql, params).fetchone()[0])  # noqa: SLF001

    def null_count(self) -> int:
        """Comparisons without an identity (no common hash), counted inside SQLite."""
        sql, params = self._where()
        return int(self.run._session.execute(  # noqa: SLF001
            "SELECT COUNT(*) - COUNT(comparisons.identity)" + sql, params).fetchone()[0])

    def values(self, columns: tuple[str, ...] = ("identity", "cov_query")):  # noqa: ANN201
        """(query_hash, subject_hash, *columns) of every comparison, ordered by query then subject, streamed from
        SQLite (no row objects: a run of 10,000 genomes has 10^8 of them)."""
        unknown = set(columns) - set(COMPARISON_COLUMNS)
        if unknown:
            msg = f"not comparison columns: {sorted(unknown)}"
            raise ValueError(msg)
        sql, params = self._where()
        return self.run._session.execute(  # noqa: SLF001
            "SELECT comparisons.query_hash, comparisons.subject_hash"
            + "".join(f", comparisons.{c}" for c in columns)
            + sql + " ORDER BY comparisons.query_hash, comparisons.subject_hash", params)

    def where_subject(self, subject_hash: str) -> "RunComparisons":  # noqa: UP037
        return RunComparisons(self.run, subject_hash)

    def __iter__(self):  # noqa: ANN204
        sql, params = self._where()
        cols = "comparisons.comparison_id, " + ", ".join("comparisons." + c for c in COMPARISON_COLUMNS)
        for row in self.run._session.execute("SELECT " + cols + sql, params):  # noqa: SLF001
            yield Comparison(*row)


class Run:
    """One all-vs-all analysis: a configuration, a set of genomes and their N x N comparisons."""

    _PERSISTED = ("status", "name", "cmdline", "df_identity", "df_cov_query", "df_aln_length",
                  "df_sim_errors", "df_hadamard")

    def __init__(  # noqa: PLR0913
        self, run_id: int | None = None, configuration_id: int | None = None, cmdline: str = "",
        fasta_directory: str = "", date: datetime.datetime | str | None = None, status: str = "",
        name: str = "", df_identity: str | None = None, df_cov_query: str | None = None,
        df_aln_length: str | None = None, df_sim_errors: str | None = None, df_hadamard: str | None = None,
        *, configuration: Configuration | None = None,
    ) -> None:
        self.__dict__["_session"] = None
        self.__dict__["_dirty"] = set()
        self.run_id = run_id
        self.configuration_id = configuration_id if configuration is None else configuration.configuration_id
        self.cmdline = cmdline
        self.fasta_directory = fasta_directory
        if isinstance(date, str):
            date = datetime.datetime.fromisoformat(date)
        self.date = date
        self.status = status
        self.name = name
        self.df_identity = df_identity
        self.df_cov_query = df_cov_query
        self.df_aln_length = df_aln_length
        self.df_sim_errors = df_sim_errors
        self.df_hadamard = df_hadamard
        self.__dict__["_configuration"] = configuration
        self._dirty.clear()

    def __setattr__(self, key: str, value: Any) -> None:
        self.__dict__[key] = value
        if key in self._PERSISTED:
            self._dirty.add(key)

    def _attach(self, session: Session) -> None:
        self.__dict__["_session"] = session
        if self.run_id is not None:
            session._runs[self.run_id] = self  # noqa: SLF001

    def _flush(self) -> None:
        if self._dirty and self._session is not None and self.run_id is not None:
            cols = sorted(self._dirty)
            self._session.execute(
                "UPDATE runs SET " + ", ".join(f"{c} = ?" for c in cols) + " WHERE run_id = ?",
                (*[getattr(self, c) for c in cols], self.run_id),
            )
            self._dirty.clear()

    # ---- relationships ----------------------------------------------------------------------
    @property
    def configuration(self) -> Configuration:
        if self._configuration is None:
            self.__dict__["_configuration"] = self._session.get_configuration(self.configuration_id)
        return self._configuration

    @property
    def fasta_hashes(self) -> QueryList:
        if self._session is None:
            return QueryList()
        rows = self._session.execute(
            "SELECT genome_hash, run_id, fasta_filename FROM runs_genomes WHERE run_id = ?", (self.run_id,)
        )
        return QueryList(RunGenomeAssociation(*r) for r in rows)

    @property
    def genomes(self) -> QueryList:
        if self._session is None:
            return QueryList()
        rows = self._session.execute(
            "SELECT genomes.genome_hash, genomes.path, genomes.length, genomes.description FROM genomes"
            " JOIN runs_genomes ON genomes.genome_hash = runs_genomes.genome_hash WHERE runs_genomes.run_id = ?",
            (self.run_id,),
        )
        return QueryList(Genome(*r) for r in rows)

    def comparisons(self) -> RunComparisons:
        """All comparison rows of this run (both genomes in the run, same configuration)."""
        return RunComparisons(self)

    # ---- cached matrices ----------------------------------------------------------------------
    def _scatter_comparisons(self, hashes: list[str], columns: tuple[str, ...]) -> "tuple[list[np.ndarray], int]":  # noqa: UP037
        """The run's comparisons as N x N float matrices (NaN where a row is missing or NULL), one per column, and
        the number of rows seen.  Row / column positions are resolved inside SQLite (a temporary md5 -> index
        table joined twice) and the rows arrive in batches that numpy scatters: no Python work per pair."""
        import numpy as np  # noqa: PLC0415

        size = len(hashes)
        out = [np.full([size, size], np.nan, float) for _ in columns]
        session = self._session
        session.execute("CREATE TEMP TABLE IF NOT EXISTS panib_matrix_index (genome_hash VARCHAR PRIMARY KEY, pos INTEGER)")
        session.execute("DELETE FROM panib_matrix_index")
        session.executemany("INSERT INTO panib_matrix_index VALUES (?, ?)", list(zip(hashes, range(size), strict=True)))
        cursor = session.execute(
            "SELECT q.pos, s.pos, " + ", ".join(f"comparisons.{c}" for c in columns) + " FROM comparisons"
            " JOIN panib_matrix_index AS q ON comparisons.query_hash = q.genome_hash"
            " JOIN panib_matrix_index AS s ON comparisons.subject_hash = s.genome_hash"
            " WHERE comparisons.configuration_id = ?", (self.configuration_id,))
        seen = 0
        while True:
            rows = cursor.fetchmany(200_000)
            if not rows:
                break
            block = np.array(rows, dtype=float)  # None -> nan
            qi, si = block[:, 0].astype(np.intp), block[:, 1].astype(np.intp)
            for m, values in zip(out, block[:, 2:].T, strict=True):
                m[qi, si] = values
            seen += len(rows)
        session.execute("DELETE FROM panib_matrix_index")
        return out, seen

    def cache_comparisons(self, *, computed: tuple | None = None) -> None:
        """Collect the N x N matrices and cache them as pandas "split" JSON (reference: db_orm.py:393-466).

        Rows = query, columns = subject, both sorted by MD5.  The caller commits.  ``computed`` =
        (sorted MD5 list, identity, cov_query) hands over matrices the caller has just recorded for a run
        that had no comparisons before, which saves reading the N^2 rows back one by one.
        """
        import numpy as np  # noqa: PLC0415
        import pandas as pd  # noqa: PLC0415

        hashes = sorted(a.genome_hash for a in self.fasta_hashes)
        size = len(hashes)
        if size * size * MATRIX_JSON_BYTES_PER_CELL > MATRIX_CACHE_MAX_BYTES:
            # SQLite refuses strings above 10^9 bytes (SQLITE_MAX_LENGTH), and the reference's cache format is
            # one JSON text per matrix: beyond ~7,000 genomes it cannot be stored.  The run stays complete -- the
            # comparisons table holds every value -- and the matrix properties below rebuild from it on demand.
            self.df_identity = self.df_cov_query = self.df_hadamard = None
            self.df_aln_length = self.df_sim_errors = None
            return
        identity = np.full([size, size], np.nan, float)
        cov_query = np.full([size, size], np.nan, float)
        aln_length = np.full([size, size], np.nan, float)
        sim_errors = np.full([size, size], np.nan, float)
        if computed is not None and list(computed[0]) == hashes:
            identity[:] = computed[1]
            cov_query[:] = computed[2]
        elif self._session is not None:
            (identity, cov_query, aln_length, sim_errors), _ = self._scatter_comparisons(
                hashes, ("identity", "cov_query", "aln_length", "sim_errors"))

        def as_json(data: "np.ndarray") -> str:  # noqa: UP037
            if size and np.isnan(data).all():
                # aln_length and sim_errors of this method: N^2 nulls, written as pandas writes them without
                # building a frame (same text; tests/test_host_boundary.py compares the two)
                names = json.dumps(hashes, separators=(",", ":"))
                row = "[" + ",".join(["null"] * size) + "]"
                return '{"columns":' + names + ',"index":' + names + ',"data":[' + ",".join([row] * size) + "]}"
            return pd.DataFrame(data=data, index=hashes, columns=hashes, dtype=float).to_json(orient="split")

        self.df_identity = as_json(identity)
        self.df_cov_query = as_json(cov_query)
        identity *= cov_query  # now hadamard
        self.df_hadamard = as_json(identity)
        self.df_aln_length = as_json(aln_length)
        self.df_sim_errors = as_json(sim_errors)

    def _matrix(self, text: str | None, *, as_float: bool, column: str | None = None) -> "DataFrame | None":  # noqa: UP037
        if not text:
            return self._matrix_from_comparisons(column) if column else None
        import pandas as pd  # noqa: PLC0415

        if as_float:
            return pd.read_json(StringIO(text), orient="split", dtype=float)
        return pd.read_json(StringIO(text), orient="split")

    def _matrix_from_comparisons(self, column: str) -> "DataFrame | None":  # noqa: UP037
        """A matrix of a COMPLETE run too large for the JSON cache (see ``cache_comparisons``), read from the
        comparisons table; None for a run that is not complete (what an empty cache means in the reference)."""
        import numpy as np  # noqa: PLC0415
        import pandas as pd  # noqa: PLC0415

        if self._session is None or self.status != "Done":
            return None
        hashes = sorted(a.genome_hash for a in self.fasta_hashes)
        size = len(hashes)
        if size * size * MATRIX_JSON_BYTES_PER_CELL <= MATRIX_CACHE_MAX_BYTES:
            return None
        cols = ("identity", "cov_query") if column == "hadamard" else (column,)
        out, seen = self._scatter_comparisons(hashes, cols)
        if seen != size * size:
            return None
        data = out[0] * out[1] if column == "hadamard" else out[0]
        return pd.DataFrame(data=data, index=hashes, columns=hashes, dtype=float)

    @property
    def identities(self) -> "DataFrame | None":  # noqa: UP037
        return self._matrix(self.df_identity, as_float=True, column="identity")

    @property
    def cov_query(self) -> "DataFrame | None":  # noqa: UP037
        return self._matrix(self.df_cov_query, as_float=True, column="cov_query")

    @property
    def aln_length(self) -> "DataFrame | None":  # noqa: UP037
        return self._matrix(self.df_aln_length, as_float=False, column="aln_length")

    @property
    def sim_errors(self) -> "DataFrame | None":  # noqa: UP037
        return self._matrix(self.df_sim_errors, as_float=False, column="sim_errors")

    @property
    def hadamard(self) -> "DataFrame | None":  # noqa: UP037
        return self._matrix(self.df_hadamard, as_float=True, column="hadamard")

    @property
    def tani(self) -> "DataFrame | None":  # noqa: UP037
        hadamard = self.hadamard
        if hadamard is None:
            return None
        return hadamard.map(lambda x: -log(x) if x else nan, na_action="ignore")

    def relabelled_matrix(self, matrix: "DataFrame", label: str = "md5") -> "DataFrame":  # noqa: UP037
        """Convert an MD5-labelled matrix of this run to "filename" or "stem" labels."""
        if label == "md5":
            return matrix
        if label == "filename":
            mapping = {a.genome_hash: a.fasta_filename for a in self.fasta_hashes}
        elif label == "stem":
            mapping = {a.genome_hash: filename_stem(a.fasta_filename) for a in self.fasta_hashes}
            if len(set(mapping.values())) < len(mapping):
                msg = "Duplicate filename stems, consider using MD5 labelling."
                raise ValueError(msg)
        else:
            msg = f"Unexpected label scheme {label!r}"
            raise ValueError(msg)
        matrix.rename(index=mapping, columns=mapping, inplace=True)  # noqa: PD002
        matrix.sort_index(axis=0, inplace=True)  # noqa: PD002
        matrix.sort_index(axis=1, inplace=True)  # noqa: PD002
        return matrix

    def __repr__(self) -> str:
        return (
            f"Run(run_id={self.run_id}, configuration_id={self.configuration_id},"
            f" cmdline={self.cmdline!r}, date={self.date!r},"
            f" status={self.status!r}, name={self.name!r}, ...)"
        )


# =============================================================================================
# module-level helpers (same names / arguments as the reference)
# =============================================================================================
def connect_to_db(logger: logging.Logger, dbpath: Path | str, *, echo: bool = False) -> Session:
    """Create/connect to the SQLite3 DB and return a session (three attempts, as the reference)."""
    import random  # noqa: PLC0415

    msg = f"Attempting to connect to {dbpath} now."
    logger.debug(msg)
    for attempt, pause in ((1, 1 + 19 * random.random()), (2, 20 + 20 * random.random()), (3, 0)):  # noqa: S311
        try:
            conn = sqlite3.connect(str(dbpath), timeout=10)
            if echo:
                conn.set_trace_callback(lambda s: logger.info(s))
            conn.executescript(SCHEMA)
            conn.commit()
            return Session(conn, dbpath)
        except sqlite3.OperationalError:  # pragma: no cover
            msg = f"Attempt {attempt}/3 failed to connect to {dbpath}"
            if attempt < 3:  # noqa: PLR2004
                logger.warning(msg)
                sleep(pause)
    log_sys_exit(logger, msg)  # pragma: no cover
    raise NotImplementedError  # pragma: no cover


def db_configuration(  # noqa: PLR0913, PLR0917
    session: Session, method: str, program: str, version: str, fragsize: int | None = None,
    mode: str | None = None, kmersize: int | None = None, minmatch: float | None = None,
    extra: str | None = None, *, create: bool = False,
) -> Configuration:
    """Return a configuration entry, adding it first if ``create`` and not already there."""
    values = (method, program, version, fragsize, mode, kmersize, minmatch, extra)
    row = session.execute(
        "SELECT configuration_id FROM configurations WHERE method IS ? AND program IS ? AND version IS ?"
        " AND fragsize IS ? AND mode IS ? AND kmersize IS ? AND minmatch IS ? AND extra IS ?", values,
    ).fetchone()
    if row is None:
        if not create:
            msg = "Requested configuration not already in DB"
            raise NoResultFound(msg)
        cur = session.execute(
            "INSERT INTO configurations (method, program, version, fragsize, mode, kmersize, minmatch, extra)"
            " VALUES (?, ?, ?, ?, ?, ?, ?, ?)", values,
        )
        session.commit()
        return Configuration(cur.lastrowid, *values)
    return Configuration(row[0], *values)


def db_genome(  # noqa: PLR0913
    logger: logging.Logger, session: Session, fasta_filename: Path | str, md5: str, *, create: bool = False,
    stats: tuple[int, bytes | None, bool] | None = None, commit: bool = True,
) -> Genome:
    """Return a genome entry, adding it first if ``create`` and not already there (trusts ``md5``).

    ``stats`` = (total bases, first title, file was gzip) when the caller has already scanned the file
    (``utils.fasta_file_stats``); otherwise the file is scanned here.  ``commit=False`` leaves the new row in
    the caller's open transaction: indexing 1,000 genomes is then one fsync instead of 1,000.
    """
    old = session.get_genome(md5)
    if old is not None:
        return old
    if not create:
        msg = "Requested genome not already in DB"
        raise NoResultFound(msg)
    name = Path(fasta_filename).name
    if stats is None:
        from pyani_plus_b200.utils import fasta_file_stats  # noqa: PLC0415

        stats = fasta_file_stats(fasta_filename)[1:]
    length, title, is_gzip = stats
    description = None if title is None else title.decode()
    if is_gzip:
        if description is None:
            msg = f"File {name} is not recognised as a FASTA record"
            log_sys_exit(logger, msg)
        if not str(fasta_filename).endswith(".gz"):
            msg = f"No .gz ending, but {name} is gzip compressed"
            log_sys_exit(logger, msg)
    elif str(fasta_filename).endswith(".gz"):
        msg = f"Has .gz ending, but {name} is NOT gzip compressed"
        log_sys_exit(logger, msg)
    old = session.get_genome(md5)
    if old is not None:
        return old  # pragma: no cover
    session.execute(
        "INSERT INTO genomes (genome_hash, path, length, description) VALUES (?, ?, ?, ?)",
        (md5, str(fasta_filename), length, description),
    )
    if commit:
        session.commit()
    return Genome(md5, str(fasta_filename), length, description)  # type: ignore[arg-type]


def add_run(  # noqa: PLR0913, PLR0917
    session: Session, configuration: Configuration, cmdline: str, fasta_directory: Path, status: str,
    name: str, date: datetime.datetime | None = None, fasta_to_hash: dict[Path, str] | None = None,
) -> Run:
    """Add and return a new run entry (makes a near-duplicate if there is a match already)."""
    date = date or datetime.datetime.now(tz=datetime.UTC)
    stored = date.replace(tzinfo=None).isoformat(sep=" ", timespec="microseconds")  # SQLAlchemy's SQLite format
    cur = session.execute(
        "INSERT INTO runs (configuration_id, cmdline, fasta_directory, date, status, name) VALUES (?, ?, ?, ?, ?, ?)",
        (configuration.configuration_id, cmdline, str(fasta_directory), stored, status, name),
    )
    run_id = cur.lastrowid
    if fasta_to_hash:
        session.executemany(
            "INSERT INTO runs_genomes (run_id, fasta_filename, genome_hash) VALUES (?, ?, ?)",
            [(run_id, Path(filename).name, md5) for filename, md5 in fasta_to_hash.items()],
        )
    session.commit()
    return session.get_run(run_id)


def load_run(
    session: Session, run_id: int | None = None, *, check_complete: bool = False, check_empty: bool = False
) -> Run:
    """Load the specified (or latest) run; optionally insist it has comparisons / is complete."""
    if run_id is None:
        row = session.execute("SELECT run_id FROM runs ORDER BY run_id DESC LIMIT 1").fetchone()
        if row is None:
            msg = "Database contains no runs."
            raise SystemExit(msg)
        run_id = row[0]
    try:
        run = session.get_run(run_id)
    except NoResultFound:
        msg = f"Database has no run-id {run_id}. Use the list-runs command for more information."
        raise SystemExit(msg) from None
    if check_complete or check_empty:
        done = run.comparisons().count()
        n = run.genomes.count()
        if not done:
            msg = f"run-id {run_id} has no comparisons"
            raise SystemExit(msg)
        if check_complete:
            if done < n**2:
                msg = f"run-id {run_id} has only {done} of {n}²={n**2} comparisons, {n**2 - done} needed"
                raise SystemExit(msg)
            if run.identities is None:
                run.cache_comparisons()
                session.commit()
    return run


def db_comparison(  # noqa: PLR0913, PLR0917
    session: Session, configuration_id: int, query_hash: str, subject_hash: str,
    identity: float | None = None, aln_length: int | None = None, sim_errors: int | None = None,
    cov_query: float | None = None, cov_subject: float | None = None,
    uname: platform.uname_result | None = None,
) -> Comparison:
    """Return a comparison entry, adding it if not already there (never alters an existing one)."""
    cols = "comparison_id, " + ", ".join(COMPARISON_COLUMNS)
    select = (f"SELECT {cols} FROM comparisons WHERE configuration_id = ? AND query_hash = ?"  # noqa: S608
              " AND subject_hash = ?")
    row = session.execute(select, (configuration_id, query_hash, subject_hash)).fetchone()
    if row is not None:
        return Comparison(*row)
    if uname is None:
        uname = platform.uname()
    values = (query_hash, subject_hash, configuration_id, identity, aln_length, sim_errors, cov_query,
              cov_subject, uname.system, uname.release, uname.machine)
    cur = session.execute(
        f"INSERT INTO comparisons ({', '.join(COMPARISON_COLUMNS)}) VALUES ({', '.join('?' * 11)})", values  # noqa: S608
    )
    session.commit()
    return Comparison(cur.lastrowid, *values)


_INSERT_OR_IGNORE = (
    f"INSERT OR IGNORE INTO comparisons ({', '.join(COMPARISON_COLUMNS)})"  # noqa: S608
    f" VALUES ({', '.join(':' + c for c in COMPARISON_COLUMNS)})"
)


def insert_comparisons_with_retries(
    logger: logging.Logger, session: Session, db_entries: list[dict[str, str | float | int | None]],
    source: str = "comparisons",
) -> bool:
    """INSERT OR IGNORE the given comparisons and commit; three attempts (reference: db_orm.py:1044-1114)."""
    import random  # noqa: PLC0415

    if not db_entries:
        session.commit()
        return True
    msg = f"Attempting to record {len(db_entries)} comparisons."
    logger.debug(msg)
    defaults = dict.fromkeys(COMPARISON_COLUMNS)
    for attempt, pause in ((1, 20 + 10 * random.random()), (2, 30 + 10 * random.random()), (3, 0)):  # noqa: S311
        try:
            session.executemany(_INSERT_OR_IGNORE, ({**defaults, **e} for e in db_entries))
            session.commit()
        except sqlite3.OperationalError:  # pragma: no cover
            msg = f"Attempt {attempt}/3 failed to record {source}"
            if attempt < 3:  # noqa: PLR2004
                logger.warning(msg)
                sleep(pause)
            else:
                logger.critical(msg)
        else:
            logger.debug("Done")
            return True
    return False  # pragma: no cover


def insert_comparison_arrays(  # noqa: PLR0913
    logger: logging.Logger, session: Session, configuration_id: int, query_hashes: list[str],
    subject_hashes: list[str], identity: Any, cov_query: Any, *, rows_per_commit: int = 2_000_000,
    rows_per_statement: int = 128,
) -> bool:
    """Array-backed ``INSERT OR IGNORE`` of a queries x subjects block (NaN -> NULL).

    Same row semantics and the same three-attempt retry as ``insert_comparisons_with_retries`` (reference:
    db_orm.py:1044-1114), without one dict -- or one Python frame, or one statement execution -- per pair:

    * each query row becomes one object array (NaN masked to None by numpy) flattened into the parameter list;
    * the seven values that are the same for every row of the call (configuration, the three NULL columns of
      this method, the uname strings) are literals of the statement, so a row binds four parameters, not eleven;
    * one statement carries ``rows_per_statement`` rows (``VALUES (...), (...), ...``; 512 parameters, below
      every SQLite build's limit), OR IGNORE still acting row by row.

    Measured per row on the same machine: 2.9 us with eleven bound parameters and one statement per row, 2.3 us
    with literals, 1.8 us with 128 rows per statement.  Committed in chunks of whole query rows, so a transient
    lock costs one chunk, not the run (what makes N = 10,000 practical).
    Returns False when a chunk could not be recorded after three attempts.
    """
    import random  # noqa: PLC0415

    import numpy as np  # noqa: PLC0415

    uname = platform.uname()
    n_q, n_s = len(query_hashes), len(subject_hashes)
    msg = f"Attempting to record {n_q * n_s} comparisons."
    logger.debug(msg)

    def literal(text: str) -> str:
        return "'" + text.replace("'", "''") + "'"

    values = {"query_hash": "?", "subject_hash": "?", "configuration_id": str(int(configuration_id)),
              "identity": "?", "aln_length": "NULL", "sim_errors": "NULL", "cov_query": "?", "cov_subject": "NULL",
              "uname_system": literal(uname.system), "uname_release": literal(uname.release),
              "uname_machine": literal(uname.machine)}
    assert set(values) == set(COMPARISON_COLUMNS)  # noqa: S101
    bound = [c for c in COMPARISON_COLUMNS if values[c] == "?"]  # order of the parameters of one row
    assert bound == ["query_hash", "subject_hash", "identity", "cov_query"]  # noqa: S101
    head = f"INSERT OR IGNORE INTO comparisons ({', '.join(COMPARISON_COLUMNS)}) VALUES "  # noqa: S608
    one_row = "(" + ", ".join(values[c] for c in COMPARISON_COLUMNS) + ")"
    per_stmt = max(1, int(rows_per_statement))
    width = len(bound) * per_stmt
    sql_full = head + ", ".join([one_row] * per_stmt)
    identity = np.asarray(identity, dtype=np.float64).reshape(n_q, n_s)
    cov_query = np.asarray(cov_query, dtype=np.float64).reshape(n_q, n_s)
    subjects = np.array(list(subject_hashes), dtype=object)

    def row_params(i: int) -> list:
        params = np.empty((n_s, len(bound)), dtype=object)
        params[:, 0] = query_hashes[i]
        params[:, 1] = subjects
        params[:, 2] = identity[i]
        params[np.isnan(identity[i]), 2] = None
        params[:, 3] = cov_query[i]
        params[np.isnan(cov_query[i]), 3] = None
        return params.ravel().tolist()

    def record(i0: int, i1: int) -> None:
        rest: list = []

        def statements():  # noqa: ANN202
            nonlocal rest
            for i in range(i0, i1):
                rest += row_params(i)
                full = len(rest) // width * width
                for j in range(0, full, width):
                    yield rest[j:j + width]
                rest = rest[full:]

        session.executemany(sql_full, statements())
        if rest:
            session.execute(head + ", ".join([one_row] * (len(rest) // len(bound))), tuple(rest))
        session.commit()

    step = max(1, rows_per_commit // max(1, n_s))
    for i0 in range(0, n_q, step):
        i1 = min(n_q, i0 + step)
        for attempt, pause in ((1, 20 + 10 * random.random()), (2, 30 + 10 * random.random()), (3, 0)):  # noqa: S311
            try:
                record(i0, i1)
            except sqlite3.OperationalError:  # pragma: no cover
                msg = f"Attempt {attempt}/3 failed to record comparisons of query rows {i0}..{i1}"
                if attempt < 3:  # noqa: PLR2004
                    logger.warning(msg)
                    sleep(pause)
                else:
                    logger.critical(msg)
                    return False
            else:
                break
    logger.debug("Done")
    return True
