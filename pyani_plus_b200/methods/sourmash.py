"""The sourmash Average Nucleotide Identity (ANI) method, computed on a B200 instead of by subprocesses.

Drop-in for ``pyani_plus/methods/sourmash.py``: same module constants and the same three functions
with the same signatures, yielded values and error behaviour --

* ``prepare_genomes(logger, run, cache)``           reference :34-84   (was: ``sourmash scripts singlesketch`` per genome)
* ``parse_sourmash_manysearch_csv(logger, file, expected_pairs)``  reference :87-144
* ``compute_sourmash_tile(logger, tool, subject_hashes, query_hashes, cache, tmp_dir)``  reference :147-206
  (was: ``sourmash sig collect`` x2 + ``sourmash scripts manysearch -m DNA -t 0``)

The cache layout (``cache/sourmash_k={k}_scaled={s}/{md5}.sig``, sourmash signature JSON) and the
``manysearch.csv`` written to ``tmp_dir`` keep the reference's formats.
"""

from __future__ import annotations

import logging
import re
from collections.abc import Iterator
from pathlib import Path

import numpy as np

from pyani_plus_b200 import db_orm, log_sys_exit, sigfile, tools, utils

SCALED = 1000
KMER_SIZE = 31  # default

# manysearch.csv is a debugging/interchange artefact; beyond this many rows it is not written
MANYSEARCH_CSV_MAX_ROWS = 1_000_000
# FASTA bytes read into host memory per sketching batch
PREPARE_BATCH_BYTES = 1 << 30

MANYSEARCH_HEADER = (
    "query_name,query_md5,match_name,containment,intersect_hashes,ksize,scaled,moltype,match_md5,jaccard,"
    "max_containment,query_containment_ani,match_containment_ani,average_containment_ani,max_containment_ani"
)

_engine = None
# sketches read or written in this process, keyed by .sig path -> (mtime_ns, parsed signature): nothing is
# parsed twice in a process, and across processes the binary side-cache (sigfile.read_side_cache) replaces
# the ~90 KB of JSON per genome
_sig_memo: dict[Path, tuple[int, dict]] = {}
SIG_MEMO_MAX = 20_000


def get_engine():  # noqa: ANN201
    """The process-wide CUDA engine (raises ``EngineError`` without a GPU: there is no CPU path)."""
    global _engine  # noqa: PLW0603
    if _engine is None:
        from pyani_plus_b200 import engine  # noqa: PLC0415

        _engine = engine.Engine()
    return _engine


def parse_scaled(logger: logging.Logger, extra: str) -> int:
    """``"scaled=1000"`` -> 1000 (the only ``extra`` form the sourmash method records)."""
    match = re.fullmatch(r"scaled=(\d+)", extra.strip())
    if not match or int(match.group(1)) < 1:
        msg = f"sourmash requires extra setting of the form scaled=<positive integer>, not {extra!r}"
        log_sys_exit(logger, msg)
    return int(match.group(1))  # type: ignore[union-attr]


def load_sketch(sig_path: Path, ksize: int | None = None) -> dict:
    """One cached sketch: from this process's memo, else the binary side-cache, else the ``.sig`` JSON.

    Returns the ``sigfile.read_sig`` dict (``md5sum`` may be "" until someone needs it: see ``sketch_md5``).
    """
    st = sig_path.stat()
    memo = _sig_memo.get(sig_path)
    if memo is not None and memo[0] == st.st_mtime_ns:
        return memo[1]
    sig = sigfile.read_side_cache(sig_path, st)
    if sig is None or (ksize is not None and sig["ksize"] != ksize):
        sig = sigfile.read_sig(sig_path, ksize=ksize)
        sigfile.write_side_cache(sig_path, sig)  # the next process skips the JSON
    if len(_sig_memo) < SIG_MEMO_MAX:
        _sig_memo[sig_path] = (st.st_mtime_ns, sig)
    return sig


def sketch_md5(sig: dict) -> str:
    """sourmash's checksum of a sketch (only the manysearch.csv artefact prints it)."""
    if not sig.get("md5sum"):
        sig["md5sum"] = sigfile.sketch_md5sum(sig["hashes"], sig["ksize"])
    return sig["md5sum"]


def sketch_entries(  # noqa: PLR0913
    logger: logging.Logger, todo: list[tuple[str, Path]], ksize: int, scaled: int, cache: Path,
) -> Iterator[tuple[str, np.ndarray]]:
    """Sketch FASTA files on the GPU and write them to the cache; yields (md5, hashes) in ``todo`` order.

    gunzip + FASTA parsing (C, GIL released) run on a thread pool a bounded window ahead; genomes are
    sketched in batches of ``PREPARE_BATCH_BYTES``; every sketch is written as sourmash ``.sig`` JSON plus
    the binary side-cache (by the same pool: decimal formatting in C, md5 and file writes release the GIL),
    and remembered for this process.
    """
    from concurrent.futures import ThreadPoolExecutor  # noqa: PLC0415

    from pyani_plus_b200 import engine  # noqa: PLC0415

    max_hash = engine.max_hash(scaled)
    pending: list[tuple[str, Path, np.ndarray]] = []
    pending_bytes = 0

    def write_one(md5: str, fasta_filename: Path, hashes: np.ndarray) -> tuple[Path, dict]:
        # formatting, md5 and file writes release the GIL: the batch's signatures are written by the pool
        sig_path = cache / f"{md5}.sig"
        sigfile.write_sig(sig_path, filename=str(fasta_filename), name=md5, ksize=ksize, max_hash=max_hash,
                          hashes=hashes)
        sig = {"name": md5, "filename": str(fasta_filename), "ksize": ksize, "seed": 42,
               "max_hash": max_hash, "md5sum": "", "hashes": hashes}
        sigfile.write_side_cache(sig_path, sig)
        return sig_path, sig

    def flush() -> Iterator[tuple[str, np.ndarray]]:
        nonlocal pending_bytes
        if pending:
            table = get_engine().sketch_genomes([recs for _, _, recs in pending], ksize, scaled)
            writes = [pool.submit(write_one, md5, fasta_filename, hashes)
                      for (md5, fasta_filename, _), hashes in zip(pending, table.to_host(), strict=True)]
            for (md5, _, _), done in zip(pending, writes, strict=True):
                sig_path, sig = done.result()
                if len(_sig_memo) < SIG_MEMO_MAX:
                    _sig_memo[sig_path] = (sig_path.stat().st_mtime_ns, sig)
                yield md5, sig["hashes"]
            pending.clear()
            pending_bytes = 0

    msg = f"Sketching {len(todo)} genomes into '{cache}'"
    logger.debug(msg)
    with ThreadPoolExecutor(max_workers=max(1, min(16, utils.available_cores()))) as pool:
        window = 4 * pool._max_workers  # noqa: SLF001
        futures = [pool.submit(utils.read_fasta_stream, path) for _, path in todo[:window]]
        for i, (md5, path) in enumerate(todo):
            if i + window < len(todo):
                futures.append(pool.submit(utils.read_fasta_stream, todo[i + window][1]))
            stream_bytes = futures[i].result()[0]
            futures[i] = None  # type: ignore[call-overload]
            pending.append((md5, path, stream_bytes))
            pending_bytes += int(stream_bytes.size)
            if pending_bytes >= PREPARE_BATCH_BYTES:
                yield from flush()
        yield from flush()


def sketches_for(  # noqa: PLR0913
    logger: logging.Logger, entries: list[tuple[str, str | Path]], ksize: int, scaled: int, cache: Path,
) -> list[np.ndarray]:
    """The sketches of ``entries`` = [(md5, FASTA path)], in order: from the cache where a signature
    exists, computed on the GPU (and cached) where not."""
    cache.mkdir(exist_ok=True)
    out: dict[str, np.ndarray] = {}
    todo = []
    for md5, path in entries:
        sig_path = cache / f"{md5}.sig"
        if sig_path.is_file():
            out[md5] = load_sketch(sig_path, ksize)["hashes"]
        else:
            todo.append((md5, Path(path)))
    out.update(sketch_entries(logger, todo, ksize, scaled, cache))
    return [out[md5] for md5, _ in entries]


def prepare_genomes(logger: logging.Logger, run: db_orm.Run, cache: Path) -> Iterator[db_orm.RunGenomeAssociation]:
    """Build the sourmash sketch signatures in the given directory.

    Will use a sub-directory ``sourmash_k={kmersize}_scaled={number}``.

    Yields the run's FASTA entries as their signatures are completed, for use with a progress bar.
    Existing signature files are never recomputed.
    """
    config = run.configuration
    if config.method != "sourmash":
        msg = f"Expected run to be for sourmash, not method {config.method}"
        log_sys_exit(logger, msg)
    if not config.kmersize:
        msg = f"sourmash requires a k-mer size, default is {KMER_SIZE}"
        log_sys_exit(logger, msg)
    if not config.extra:
        msg = f"sourmash requires extra setting, default is scaled={SCALED}"
        log_sys_exit(logger, msg)
    tools.get_sourmash()
    if not cache.is_dir():
        msg = f"Cache directory '{cache}' does not exist"
        raise ValueError(msg)
    scaled = parse_scaled(logger, config.extra)
    ksize = int(config.kmersize)
    cache = cache / f"sourmash_k={config.kmersize}_{config.extra}"
    msg = f"Preparing sourmash signatures in '{cache}'"
    logger.debug(msg)
    cache.mkdir(exist_ok=True)
    fasta_dir = Path(run.fasta_directory)

    todo = {}
    for entry in run.fasta_hashes:
        if (cache / f"{entry.genome_hash}.sig").is_file():
            yield entry
        else:
            todo[entry.genome_hash] = entry
    work = [(md5, fasta_dir / e.fasta_filename) for md5, e in todo.items()]
    for md5, _ in sketch_entries(logger, work, ksize, scaled, cache):
        yield todo[md5]


def parse_sourmash_manysearch_csv(
    logger: logging.Logger,
    manysearch_file: Path,
    expected_pairs: set[tuple[str, str]],
) -> Iterator[tuple[str, str, float | None, float | None]]:
    """Parse sourmash-plugin-branchwater manysearch CSV output.

    Returns tuples of (query_hash, subject_hash, query-containment ANI estimate, max-containment
    ANI estimate).  Any pairs not in the file are inferred to be failed alignments: with reporting
    threshold zero that only happens when two sketches share no hash at all.
    """
    with manysearch_file.open() as handle:
        header_line = handle.readline().rstrip("\n")
        headers = header_line.split(",")
        try:
            # column order differs between branchwater commands / versions, so go by name
            col_query = headers.index("query_name")
            col_subject = headers.index("match_name")
            col_query_cont = headers.index("query_containment_ani")
            col_max_cont = headers.index("max_containment_ani")
        except ValueError:
            msg = f"Missing expected fields in sourmash manysearch header, found: {header_line!r}"
            log_sys_exit(logger, msg)
        for raw in handle:
            line = raw.rstrip("\n")
            if not line:
                continue
            values = line.split(",")
            query_hash, subject_hash = values[col_query], values[col_subject]
            if query_hash == subject_hash and values[col_max_cont] != "1.0":
                msg = f"Expected sourmash manysearch {query_hash} vs self to be one, not {values[col_max_cont]!r}"
                raise ValueError(msg)
            if (query_hash, subject_hash) in expected_pairs:
                expected_pairs.remove((query_hash, subject_hash))
            else:
                msg = f"Did not expect {query_hash} vs {subject_hash} in {manysearch_file.name}"
                log_sys_exit(logger, msg)
            yield query_hash, subject_hash, float(values[col_query_cont]), float(values[col_max_cont])
    # even if the file was empty (bar the header), remaining pairs are failed alignments
    for query_hash, subject_hash in expected_pairs:
        yield query_hash, subject_hash, None, None


def _fmt(x: float) -> str:
    """Shortest round-trip decimal without exponent (how Rust's ``{}`` prints an f64)."""
    return np.format_float_positional(x, unique=True, trim="0")


def write_manysearch_csv(  # noqa: PLR0913
    path: Path, queries: list[str], subjects: list[str], q_md5: list[str], s_md5: list[str],
    q_counts: np.ndarray, s_counts: np.ndarray, ov: np.ndarray, ksize: int, scaled: int,
) -> int:
    """Write the branchwater ``manysearch`` table for the computed block; returns the rows written.

    Rows exist only for pairs with at least one common hash (SURVEY.md 8c convention 13).
    """
    rows = 0
    with path.open("w") as handle:
        handle.write(MANYSEARCH_HEADER + "\n")
        for i, q in enumerate(queries):
            nq = int(q_counts[i])
            for j in np.flatnonzero(ov[i]):
                n_ov, ns = int(ov[i, j]), int(s_counts[j])
                cq, cs = n_ov / nq, n_ov / ns
                qani, mani = _ani(cq, ksize), _ani(cs, ksize)
                handle.write(
                    f"{q},{q_md5[i]},{subjects[j]},{_fmt(cq)},{n_ov},{ksize},{scaled},DNA,{s_md5[j]},"
                    f"{_fmt(n_ov / (nq + ns - n_ov))},{_fmt(max(cq, cs))},{_fmt(qani)},{_fmt(mani)},"
                    f"{_fmt((qani + mani) / 2.0)},{_fmt(max(qani, mani))}\n"
                )
                rows += 1
    return rows


def _ani(containment: float, ksize: int) -> float:
    """sourmash ``ani_from_containment`` (only used for the optional CSV artefact)."""
    if containment == 0.0:
        return 0.0
    if containment == 1.0:
        return 1.0
    return 1.0 - (1.0 - containment ** (1.0 / ksize))


def tile_arrays(  # noqa: PLR0913
    logger: logging.Logger,
    subject_hashes: set[str],
    query_hashes: set[str],
    cache: Path,
    tmp_dir: Path | None,
) -> tuple[list[str], list[str], np.ndarray, np.ndarray, np.ndarray]:
    """GPU part of ``compute_sourmash_tile``: (queries, subjects, ov, identity, cov_query) as arrays.

    ``queries`` / ``subjects`` are the sorted MD5 lists labelling the rows / columns; ``ov`` holds the
    intersection sizes, ``identity`` / ``cov_query`` the max- and query-containment ANI with NaN
    where branchwater would print no row.
    """
    m = re.fullmatch(r"sourmash_k=(\d+)_scaled=(\d+)", cache.name)
    queries, subjects = sorted(query_hashes), sorted(subject_hashes)
    loaded: dict[str, dict] = {}
    for md5 in sorted(set(queries) | set(subjects)):
        sig_path = cache / f"{md5}.sig"
        if not sig_path.is_file():
            msg = f"Missing sourmash signature file '{sig_path}'"
            log_sys_exit(logger, msg)
        loaded[md5] = load_sketch(sig_path, int(m.group(1)) if m else None)
    if not loaded:
        empty = np.zeros((0, 0))
        return queries, subjects, empty.astype(np.uint32), empty, empty
    ksizes = {s["ksize"] for s in loaded.values()}
    max_hashes = {s["max_hash"] for s in loaded.values()}
    if len(ksizes) != 1 or len(max_hashes) != 1:
        msg = f"Inconsistent sourmash signatures under '{cache}': ksize {sorted(ksizes)}, max_hash {sorted(max_hashes)}"
        log_sys_exit(logger, msg)
    ksize = ksizes.pop()
    from pyani_plus_b200 import engine  # noqa: PLC0415

    scaled = int(m.group(2)) if m else 0
    if not scaled or engine.max_hash(scaled) != next(iter(max_hashes)):
        # recover scaled from max_hash (sourmash: scaled = round(2^64 / max_hash))
        scaled = max(1, round(2**64 / max(1, next(iter(max_hashes)))))
        if engine.max_hash(scaled) != next(iter(max_hashes)):
            msg = f"Cannot determine scaled for max_hash={next(iter(max_hashes))}"
            log_sys_exit(logger, msg)

    eng = get_engine()
    q_table = eng.table_from_host([loaded[h]["hashes"] for h in queries], ksize, scaled)
    if queries == subjects:
        s_table = q_table
        ov = eng.intersect(q_table).cpu().numpy()
    else:
        s_table = eng.table_from_host([loaded[h]["hashes"] for h in subjects], ksize, scaled)
        ov = eng.intersect(q_table, s_table).cpu().numpy()
    ov = ov.astype(np.uint32, copy=False)
    q_counts = q_table.counts.cpu().numpy()
    s_counts = s_table.counts.cpu().numpy()
    # host finalisation with libm pow: the floats are the ones branchwater prints
    identity, cov_query = engine.ani_host(ov, q_counts, s_counts, ksize)
    if tmp_dir is not None and len(queries) * len(subjects) <= MANYSEARCH_CSV_MAX_ROWS:
        write_manysearch_csv(
            tmp_dir / "manysearch.csv", queries, subjects, [sketch_md5(loaded[h]) for h in queries],
            [sketch_md5(loaded[h]) for h in subjects], q_counts, s_counts, ov, ksize, scaled,
        )
    column_of = {s: j for j, s in enumerate(subjects)}
    for i, q in enumerate(queries):  # self-vs-self must be one
        j = column_of.get(q)
        if j is not None and ov[i, j] and identity[i, j] != 1.0:
            msg = f"Expected sourmash manysearch {queries[i]} vs self to be one, not {identity[i, j]!r}"
            raise ValueError(msg)
    return queries, subjects, ov, identity, cov_query


def compute_sourmash_tile(  # noqa: PLR0913, PLR0917
    logger: logging.Logger,
    tool: tools.ExternalToolData,  # noqa: ARG001
    subject_hashes: set[str],
    query_hashes: set[str],
    cache: Path,
    tmp_dir: Path,
) -> Iterator[tuple[str, str, float | None, float | None]]:
    """Intersect the cached sketches of queries x subjects on the GPU and return pairwise ANI values.

    Yields (query_hash, subject_hash, query-containment ANI, max-containment ANI) for EVERY ordered
    pair; pairs without a common hash (no manysearch row) carry ``None, None``.
    """
    if not cache.is_dir():
        msg = f"Given cache directory '{cache}' does not exist"
        raise ValueError(msg)
    query_sig_list = tmp_dir / "query_sigs.csv"
    subject_sig_list = tmp_dir / "subject_sigs.csv"
    for csv, sigs in ((query_sig_list, query_hashes), (subject_sig_list, subject_hashes)):
        if csv.is_file():
            msg = f"Race condition? Replacing intermediate file '{csv}'"
            logger.warning(msg)
            csv.unlink()
        csv.write_text("internal_location\n" + "".join(f"{cache / (_ + '.sig')}\n" for _ in sorted(sigs)))
    queries, subjects, ov, identity, cov_query = tile_arrays(logger, subject_hashes, query_hashes, cache, tmp_dir)
    missing: list[tuple[str, str]] = []
    for i, q in enumerate(queries):
        id_row, cov_row, ov_row = identity[i], cov_query[i], ov[i]
        for j, s in enumerate(subjects):
            if ov_row[j]:
                yield q, s, float(cov_row[j]), float(id_row[j])
            else:
                missing.append((q, s))
    # pairs without a manysearch row are failed alignments (no common k-mer hashes)
    for q, s in missing:
        yield q, s, None, None
