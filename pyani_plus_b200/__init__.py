"""B200-native drop-in for the ``sourmash`` method of pyani-plus.

Only the one hot path is implemented (FracMinHash sketching + all-vs-all sketch intersection ->
max-containment ANI), behind the reference's own interfaces for that path:

* ``pyani_plus_b200.methods.sourmash``  <->  ``pyani_plus/methods/sourmash.py``
* ``pyani_plus_b200.private_cli``       <->  ``pyani_plus/private_cli.py`` (prepare, compute-column, JSON)
* ``pyani_plus_b200.db_orm``            <->  ``pyani_plus/db_orm.py`` (same SQLite schema, stdlib sqlite3)
* ``pyani_plus_b200.public_cli``        <->  ``pyani_plus/public_cli.py`` (``sourmash`` and ``resume``)

The arithmetic runs in hand-written sm_100a CUDA kernels (``csrc/``) reached through the C ABI
declared in ``include/panib200.h``; there is no CPU fallback.

This module mirrors the few package-level helpers of ``pyani_plus/__init__.py`` the path uses
(``setup_logger`` :61-117, ``log_sys_exit`` :120-126, ``FASTA_EXTENSIONS`` :48).
"""

from __future__ import annotations

import logging
import sys
from pathlib import Path

__version__ = "0.1.0"

LOG_FILE = Path("pyani-plus.log")
LOG_FILE_DYNAMIC = Path("--")  # internal use only, not exposed in CLI
FASTA_EXTENSIONS = {".fasta", ".fas", ".fna", ".fa"}  # also with .gz appended


_FILE_FORMAT = "%(asctime)s %(levelname)9s %(filename)21s:%(lineno)-3s | %(message)s"


def _console_handler(level: int, *, plain: bool) -> logging.Handler:
    """Terminal handler: plain stderr stream for workers, rich (stdout, with markup) for the public CLI."""
    if plain:
        handler: logging.Handler = logging.StreamHandler()
        handler.setLevel(level)
        return handler
    from rich.logging import RichHandler  # noqa: PLC0415

    return RichHandler(level=level, markup=True, omit_repeated_times=False, show_path=False,
                       rich_tracebacks=True, tracebacks_suppress=["click"])


def setup_logger(
    log_file: Path | None, *, terminal_level: int = logging.INFO, plain: bool = False
) -> logging.Logger:
    """Package logger with a console handler and, optionally, a DEBUG-level log file.

    Same contract as the reference's ``setup_logger`` (``pyani_plus/__init__.py:61-117``): ``None`` or
    ``Path("-")`` means no file; calling it again replaces the previous handlers; the file always
    records DEBUG while the terminal shows ``terminal_level`` and up.
    """
    if log_file == LOG_FILE_DYNAMIC:
        sys.exit("ERROR: Internal flag value for dynamic log setting unresolved")
    logger = logging.getLogger(__package__)
    logger.handlers.clear()  # repeated set-up must not duplicate output
    logger.setLevel(min(logging.DEBUG, terminal_level))
    logger.addHandler(_console_handler(terminal_level, plain=plain))
    wants_file = log_file is not None and log_file != Path("-")
    if wants_file:
        to_file = logging.FileHandler(log_file, mode="a")
        to_file.setLevel(logging.DEBUG)
        to_file.setFormatter(logging.Formatter(fmt=_FILE_FORMAT, datefmt="%Y-%m-%d %H:%M:%S"))
        logger.addHandler(to_file)
        logger.info("Logging to '%s'", log_file)  # shown on the terminal too
    else:
        logger.debug("Currently not logging to file.")
    return logger


def log_sys_exit(logger: logging.Logger, msg: str) -> None:
    """Log CRITICAL level message, then exit with that message (reference: __init__.py:120-126)."""
    logger.critical(msg)
    sys.exit(msg)
