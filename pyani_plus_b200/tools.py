"""Tool discovery for the sourmash path.

The reference finds the external ``sourmash`` binary and parses ``sourmash -v``
(``pyani_plus/tools.py:381-405``) and records ``(exe_path.stem, version)`` in the run's
``Configuration``; resuming a run requires both to match (``private_cli.py:191-223``).
Here there is no external binary: the "tool" is the in-tree CUDA engine, reported under its own
program name and version so that caches and databases made by real sourmash are never silently
mixed with ours.
"""

from __future__ import annotations

from pathlib import Path
from typing import NamedTuple


class ExternalToolData(NamedTuple):
    """Convenience struct for tool path and version information (reference: tools.py:36-40)."""

    exe_path: Path
    version: str


def get_sourmash(cmd: str | Path | None = None) -> ExternalToolData:  # noqa: ARG001
    """Return the engine standing in for ``sourmash`` + ``sourmash_plugin_branchwater``.

    Fails loudly (``EngineError``) if ``libpanib200.so`` has not been built.
    """
    from pyani_plus_b200 import engine  # noqa: PLC0415

    lib = engine.load_library()
    version = lib.panib_version().decode().split()[0]
    return ExternalToolData(exe_path=Path(engine.PROGRAM), version=version)
