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
