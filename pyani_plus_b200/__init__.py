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


def setup_logger(
    log_file: Path | None, *, terminal_level: int = logging.INFO, plain: bool = False
) -> logging.Logger:
    """Return a file-based logger alongside a console logger (reference: __init__.py:61-117).

    ``Path("-")`` or ``None`` means no log file.  The file handler is always at DEBUG level.
    """
    if log_file == LOG_FILE_DYNAMIC:
        sys.exit("ERROR: Internal flag value for dynamic log setting unresolved")
    logger = logging.getLogger(f"{__package__}")
    min_level = min(logging.DEBUG, terminal_level)
    logger.setLevel(min_level)
    if logger.hasHandlers():
        logger.handlers.clear()
    logging.basicConfig(level=min_level, format="%(message)s", datefmt="[%X]", handlers=[])

    console_handler: logging.Handler
    if plain:
        console_handler = logging.StreamHandler()
        console_handler.setLevel(terminal_level)
    else:
        from rich.logging import RichHandler  # noqa: PLC0415

        console_handler = RichHandler(
            level=terminal_level,
            markup=True,
            omit_repeated_times=False,
            show_path=False,
            rich_tracebacks=True,
            tracebacks_suppress=["click"],
        )
    logger.addHandler(console_handler)

    if log_file and log_file != Path("-"):
        file_handler = logging.FileHandler(log_file, mode="a")
        file_handler.setLevel(logging.DEBUG)
        file_handler.setFormatter(
            logging.Formatter(
                fmt="%(asctime)s %(levelname)9s %(filename)21s:%(lineno)-3s | %(message)s",
                datefmt="%Y-%m-%d %H:%M:%S",
            )
        )
        logger.addHandler(file_handler)
        msg = f"Logging to '{log_file}'"
        logger.info(msg)
    else:
        logger.debug("Currently not logging to file.")
    return logger


def log_sys_exit(logger: logging.Logger, msg: str) -> None:
    """Log CRITICAL level message, then exit with that message (reference: __init__.py:120-126)."""
    logger.critical(msg)
    sys.exit(msg)
