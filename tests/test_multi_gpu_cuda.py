"""Two (or more) real GPUs: sliced K1, fused / NCCL exchange of sketches, sharded K2 == single GPU."""

from __future__ import annotations

import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu


def test_two_gpu_pipeline_matches_single_gpu() -> None:
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    worker = Path(__file__).with_name("multi_gpu_worker.py")
    proc = subprocess.run(  # noqa: S603
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
         "--master-addr", "127.0.0.1", "--master-port", "29533", str(worker)],
        capture_output=True, text=True, timeout=600, check=False,
    )
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    assert "MULTI_GPU_OK" in proc.stdout, proc.stdout[-2000:]


def test_cli_under_torchrun_equals_single_process(tmp_path: Path) -> None:
    """`pyani-plus sourmash` launched by torchrun on 2 GPUs (the product's multi-GPU entry: rank 0 owns the
    database, every rank sketches a slice, one exchange, sharded K2) records the same rows and writes the
    same signatures as the single-process run -- genomes of unequal length, one gzipped."""
    import gzip
    import sqlite3

    import torch

    from oracle import oracle

    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    fasta = tmp_path / "genomes"
    fasta.mkdir()
    for g in range(7):
        seq = oracle.synth_genome(20261017, g, 120_000 + 90_000 * g)
        text = b">genome_%d synthetic\n" % g + b"\n".join(seq[i: i + 80] for i in range(0, len(seq), 80)) + b"\n"
        if g == 3:
            (fasta / f"g{g}.fna.gz").write_bytes(gzip.compress(text))
        else:
            (fasta / f"g{g}.fna").write_bytes(text)
    root = Path(__file__).resolve().parent.parent
    # (no --log here: torchrun's own parser would claim it as an abbreviation of --log-dir)
    common = ["sourmash", str(fasta), "--create-db", "--scaled", "100"]
    (tmp_path / "c1").mkdir()
    one = subprocess.run(  # noqa: S603
        [sys.executable, "-m", "pyani_plus_b200.public_cli", *common, "-d", str(tmp_path / "one.db"), "--cache",
         str(tmp_path / "c1")], capture_output=True, text=True, timeout=600, check=False, cwd=root)
    assert one.returncode == 0, one.stdout[-2000:] + one.stderr[-2000:]
    (tmp_path / "c2").mkdir()
    two = subprocess.run(  # noqa: S603
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
         "127.0.0.1", "--master-port", "29541", "-m", "pyani_plus_b200.public_cli", *common, "-d",
         str(tmp_path / "two.db"), "--cache", str(tmp_path / "c2")],
        capture_output=True, text=True, timeout=600, check=False, cwd=root)
    assert two.returncode == 0, two.stdout[-3000:] + two.stderr[-3000:]

    def rows(db: Path) -> list[tuple]:
        with sqlite3.connect(db) as conn:
            return sorted(conn.execute("SELECT query_hash, subject_hash, identity, cov_query FROM comparisons"))

    a, b = rows(tmp_path / "one.db"), rows(tmp_path / "two.db")
    assert len(a) == 49 and a == b
    sigs1 = {p.name: p.read_bytes() for p in (tmp_path / "c1").rglob("*.sig")}
    sigs2 = {p.name: p.read_bytes() for p in (tmp_path / "c2").rglob("*.sig")}
    assert len(sigs1) == 7 and set(sigs1) == set(sigs2)
    for name, data in sigs1.items():  # the "filename" field differs only if the paths do: same FASTA dir here
        assert data == sigs2[name], name
    with sqlite3.connect(tmp_path / "two.db") as conn:
        status, ident = conn.execute("SELECT status, df_identity FROM runs").fetchone()
    assert status == "Done" and ident
