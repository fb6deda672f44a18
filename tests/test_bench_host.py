"""Host logic of bench.py that needs no GPU: the config object both arms print, the oracle-pinned workload
checksums, and the rule that ties the ncu-derived roofline fields to the sources they were measured from."""

from __future__ import annotations

import json
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import bench  # noqa: E402


def test_both_arms_print_the_same_config() -> None:
    """The GPU arm and the reference arm build ``config`` from (workload, N) only (same_config)."""
    for workload in ("config2", "config3", "config4", "config5"):
        for n in (1, 2, 8):
            assert bench.config_dict(workload, n) == bench.config_dict(workload, n)
        cfg = bench.config_dict(workload, 1)
        n_genomes = bench.WORKLOADS[workload][0]
        assert cfg["pairs"] == n_genomes * (n_genomes - 1) // 2
        assert "model" not in cfg and cfg["workload"].startswith("configs[")


def test_default_workload_is_baseline_configs2() -> None:
    """The default bench workload is the configuration BASELINE.json names for 1/2/4/8 GPUs (configs[2])."""
    baseline = json.loads((ROOT / "BASELINE.json").read_text())
    assert "1,000 synthetic 5 Mb genomes" in baseline["configs"][2]
    n, length, k, scaled, desc = bench.WORKLOADS["config3"]
    assert (n, length, k, scaled) == (1000, 5_000_000, 31, 1000) and desc.startswith("configs[2]")
    import re

    assert re.search(r'"--workload".*default="config3"', (ROOT / "bench.py").read_text())


def test_every_baseline_workload_has_a_pinned_checksum() -> None:
    """tools/oracle_checksums.py ran the CPU oracle over configs[1]-[4]; bench.py exits non-zero on a mismatch."""
    for workload in ("config2", "config3", "config4", "config5"):
        want = bench.expected_checksum(workload)
        assert want is not None, workload
        assert set(want) >= {"ov_weighted_sum", "hash_sum", "sketch_total"}


def test_kernel_source_hashes_cover_the_include_graph() -> None:
    csrc = ROOT / "pyani_plus_b200" / "csrc"
    for kernel, files in bench.KERNEL_SOURCES.items():
        for name in files:
            assert (csrc / name).is_file(), (kernel, name)
        unit = next(f for f in files if f.endswith(".cu"))
        included = {ln.split('"')[1] for ln in (csrc / unit).read_text().splitlines()
                    if ln.startswith('#include "') and not ln.split('"')[1].startswith("..")}
        assert included <= set(files), f"{kernel}: {unit} includes {included - set(files)}"
    hashes = {k: bench.kernel_source_sha(k) for k in bench.KERNEL_SOURCES}
    assert len(set(hashes.values())) == len(hashes)
    assert bench.kernel_source_sha() not in hashes.values()


def test_ncu_fields_are_dropped_when_the_kernel_changed(monkeypatch: pytest.MonkeyPatch) -> None:
    """A capture is evidence only for the sources it measured: with a different hash bench prints nulls."""
    real = bench.kernel_source_sha
    monkeypatch.setattr(bench, "kernel_source_sha", lambda kernel=None: "0" * 16)
    assert bench.ncu_capture("config3", "k1") is None
    assert bench.ncu_capture("config3", "k2_index") is None
    monkeypatch.setattr(bench, "kernel_source_sha", real)
    assert bench.ncu_capture("config3", "no_such_kernel") is None
    assert bench.ncu_capture("no_such_workload", "k1") is None


def test_committed_ncu_captures_match_this_tree() -> None:
    """The captures under profiles/ were taken from the kernels as they are in this tree (if a kernel source
    changes, re-run tools/profile_r2.sh + tools/ncu_to_json.py, or the roofline loses its measured traffic)."""
    for kernel in ("k1", "k2_index"):
        cap = bench.ncu_capture("config3", kernel)
        assert cap is not None, f"profiles/ncu_r*.json is stale for {kernel}"
        assert cap["dram_bytes"] > 0 and cap["duration_ms"] > 0
    k1 = bench.ncu_capture("config3", "k1")
    # the numbers DESIGN.md quotes: ~116 warp-instructions per k-mer, DRAM traffic ~ the algorithmic bytes
    per_kmer = k1["inst_executed"] * 32 / k1["bases"]
    assert 100 < per_kmer < 125
    assert 0.9 < k1["dram_bytes"] / (0.383 * k1["bases"]) < 1.2
