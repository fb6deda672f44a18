"""Parity at BASELINE.json's full sizes (SURVEY.md 8d "Parity at scale").

config 3 (1000 x 5 Mb): every sketch and the complete 1000 x 1000 count matrix against the oracle.
config 5 (2000 x 5 Mb, scaled=100) and config 4 (10 000 x 5 Mb): the GPU runs the FULL workload; the
oracle checks blocks of genomes spread over the id range (all their hashes, their sub-matrix of the
full count matrix) and the whole matrix must satisfy the size-independent properties (symmetric,
diagonal = sketch size, ov <= min size).  ``tools/parity_at_scale.py`` runs the same check with any
sample size; profiles/r01_parity_at_scale.md records the complete runs.
"""

from __future__ import annotations

import pytest

from tools.parity_at_scale import check_workload

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize(("workload", "sample", "k2"), [("config3", 0, "probe"), ("config3", 0, "index"),
                                                        ("config5", 208, "auto"), ("config4", 256, "auto")])
def test_baseline_config_against_oracle(workload: str, sample: int, k2: str) -> None:
    import torch

    if workload == "config4" and torch.cuda.mem_get_info(0)[1] < 60 << 30:
        pytest.skip("needs a 64 GB+ GPU")
    report = check_workload(workload, sample, k2_method=k2)
    assert k2 == "auto" or report["k2_method"] == k2
    assert report["sketches_with_mismatch"] == 0, report
    assert report["count_cells_with_mismatch"] == 0, report
    assert report["ani_device_vs_host_max_abs_err"] <= 1e-12, report
    assert report["ani_host_rows_not_equal_oracle"] == 0, report
    # whole-result checksum == the ORACLE's over the complete workload (tests/golden/workload_checksums.json)
    assert report["oracle_checksum_of_complete_workload"] is not None and report["checksum_ok"], report
    assert report["ok"], report
    assert report["whole_matrix"]["null_pairs"] > 0  # distant descendants share no hash (NULL path)
