"""pytest configuration: markers, paths and shared fixtures."""

from __future__ import annotations

import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config: pytest.Config) -> None:
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden() -> Path:
    return GOLDEN


@pytest.fixture(scope="session")
def input_genomes_tiny() -> Path:
    """Same name as the reference's fixture (tests/conftest.py there): the viral example."""
    return GOLDEN / "viral_example"


@pytest.fixture(scope="session")
def input_genomes_bad_alignments() -> Path:
    return GOLDEN / "bad_alignments"


@pytest.fixture(scope="session")
def input_bacteria() -> Path:
    return GOLDEN / "bacterial_example"
