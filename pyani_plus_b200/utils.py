"""Host helpers that define genome identity and length for the sourmash path.

Mirrors the parts of ``pyani_plus/utils.py`` the path depends on -- same names, arguments and
error behaviour -- written from the documented behaviour, not copied:

* ``fasta_bytes_iterator``  (utils.py:40-90)   record parsing; sequences with all whitespace removed
* ``file_md5sum``           (utils.py:142-196) md5 of the *decompressed* file contents
* ``check_db`` / ``check_fasta`` (utils.py:216-242), ``filename_stem`` (:93-107), ``available_cores`` (:199-214)
"""

from __future__ import annotations

import gzip
import hashlib
import logging
import os
from collections.abc import Iterator
from pathlib import Path
from typing import IO

from pyani_plus_b200 import FASTA_EXTENSIONS, log_sys_exit

_GT = ord(">")
_WHITESPACE = b" \t\r\n"


def fasta_bytes_iterator(handle: IO[bytes] | gzip.GzipFile) -> Iterator[tuple[bytes, bytes]]:
    """Yield (title without ">", sequence) per FASTA record from a binary handle.

    Text before the first ``>`` line is skipped; the title is right-stripped; the sequence has
    every space, tab, CR and LF removed (so embedded CRs of mangled files disappear too).
    """
    if handle.read(0) != b"":
        msg = "Function fasta_bytes_iterator requires a handle in binary mode"
        raise ValueError(msg)
    title: bytes | None = None
    parts: list[bytes] = []
    for line in handle:
        if line[0] == _GT:
            if title is not None:
                yield title, b"".join(parts).translate(None, _WHITESPACE)
            title = line[1:].rstrip()
            parts = []
        elif title is not None:
            parts.append(line)
    if title is not None:
        yield title, b"".join(parts).translate(None, _WHITESPACE)


def open_maybe_gzip(filename: Path | str) -> IO[bytes]:
    """Binary handle on the decompressed contents (gzip detected by its magic bytes)."""
    handle = Path(filename).open("rb")  # noqa: SIM115
    magic = handle.read(2)
    handle.seek(0)
    if magic == b"\x1f\x8b":
        return gzip.GzipFile(fileobj=handle, mode="rb")  # type: ignore[return-value]
    return handle


def read_fasta_records(filename: Path | str) -> list[tuple[bytes, bytes]]:
    """All (title, sequence) records of a plain or gzipped FASTA file."""
    with open_maybe_gzip(filename) as handle:
        return list(fasta_bytes_iterator(handle))


def read_fasta_stream(filename: Path | str):  # noqa: ANN201
    """Read a plain or gzipped FASTA file into base-stream form with the C parser of libpanib200.

    Returns ``(stream uint8 array, n_records, total_bases, first title or None)``; the records are
    the ones ``fasta_bytes_iterator`` yields, joined by one invalid ``N``.  zlib, the C parser and
    ctypes all release the GIL, so files can be read by a thread pool.
    """
    from pyani_plus_b200 import engine  # noqa: PLC0415

    raw = Path(filename).read_bytes()
    if raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    return engine.fasta_to_stream(raw)


def fasta_file_stats(filename: Path | str) -> tuple[str, int, bytes | None, bool]:
    """One pass over a FASTA file: (md5 of decompressed bytes, total bases, first title, was gzip).

    What ``file_md5sum`` plus the length / description scan of ``db_genome`` compute, from a single
    read; hashlib, zlib and the C parser release the GIL, so a thread pool scales over files.
    """
    from pyani_plus_b200 import engine  # noqa: PLC0415

    fname = Path(filename)
    try:
        raw = fname.read_bytes()
    except FileNotFoundError:
        msg = f"Input {fname} is a broken symlink" if fname.is_symlink() else f"Input {fname} not found"
        raise ValueError(msg) from None
    is_gzip = raw[:2] == b"\x1f\x8b"
    if is_gzip:
        raw = gzip.decompress(raw)
    md5 = hashlib.md5(raw).hexdigest()  # noqa: S324
    lib = engine.load_library()
    import ctypes  # noqa: PLC0415

    out4 = (ctypes.c_int64 * 4)()
    lib.panib_fasta_to_stream(raw, len(raw), None, 0, out4)
    title = raw[out4[2]: out4[2] + out4[3]] if out4[2] >= 0 else None
    return md5, int(out4[1]), title, is_gzip


def filename_stem(filename: str) -> str:
    """Basename without the FASTA extension, also dropping a ``.gz`` suffix.

    >>> filename_stem("/path/example.fna")
    'example'
    >>> filename_stem("relative/path/example.fna.gz")
    'example'
    """
    name = filename.rsplit("/", 1)[-1]
    if name.endswith(".gz"):
        name = name[:-3]
    return Path(name).stem


def str_md5sum(text: str, encoding: str = "ascii") -> str:
    """MD5 hex digest of a string."""
    return hashlib.md5(text.encode(encoding)).hexdigest()  # noqa: S324


def file_md5sum(filename: Path | str) -> str:
    """MD5 hex digest of the file contents; for gzip files, of the decompressed contents.

    Raises ``ValueError`` ("Input X not found" / "Input X is a broken symlink") for a missing file.
    """
    fname = Path(filename)
    digest = hashlib.md5()  # noqa: S324
    try:
        with open_maybe_gzip(fname) as handle:
            while chunk := handle.read(1 << 20):
                digest.update(chunk)
    except FileNotFoundError:
        msg = f"Input {fname} is a broken symlink" if fname.is_symlink() else f"Input {fname} not found"
        raise ValueError(msg) from None
    return digest.hexdigest()


def available_cores() -> int:
    """How many CPU cores/threads are available to this process."""
    try:
        return len(os.sched_getaffinity(0))  # type: ignore[attr-defined]
    except AttributeError:  # pragma: no cover
        cpus = os.cpu_count()
        if not cpus:
            msg = "Cannot determine CPU count"
            raise RuntimeError(msg) from None
        return cpus


def check_db(logger: logging.Logger, database: Path | str, create_db: bool) -> None:  # noqa: FBT001
    """Check DB exists, or using create_db=True."""
    msg = f"Checking DB argument '{database}'"
    logger.debug(msg)
    if database != ":memory:" and not create_db and not Path(database).is_file():
        msg = f"Database {database} does not exist, but not using --create-db"
        log_sys_exit(logger, msg)


def check_fasta(logger: logging.Logger, fasta: Path) -> list[Path]:
    """Check fasta is a directory and return list of FASTA files in it."""
    msg = f"Checking FASTA argument '{fasta}'"
    logger.debug(msg)
    if not fasta.is_dir():
        msg = f"FASTA input {fasta} is not a directory"
        log_sys_exit(logger, msg)
    names: list[Path] = []
    for ext in FASTA_EXTENSIONS:
        names.extend(fasta.glob("*" + ext))
        names.extend(fasta.glob("*" + ext + ".gz"))
    if not names:
        msg = f"No FASTA input genomes under {fasta} with extensions {', '.join(FASTA_EXTENSIONS)}"
        log_sys_exit(logger, msg)
    return names
