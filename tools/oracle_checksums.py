#!/usr/bin/env python
"""Pin the BASELINE workloads' results with the CPU oracle (test infrastructure, not product code).

    python tools/oracle_checksums.py config2 config3 config5 config4

For every named workload of bench.py the ORACLE sketches all genomes of the synthetic recipe
(SURVEY.md 8d), intersects ALL pairs and reduces the result to the three numbers bench.py prints as
``result_checksum`` (position-weighted wrapping sum of the complete count matrix, wrapping sum of every
sketch hash, number of hashes).  They are merged into tests/golden/workload_checksums.json; bench.py
exits non-zero when a GPU run (any --gpus N, either K2 form) disagrees, and the GPU tests assert them.
Config 4 (10,000 genomes, 5*10^7 pairs) takes a few minutes on 8 cores.
"""

from __future__ import annotations

import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from bench import SEED, WORKLOADS, checksum_weights  # noqa: E402
from oracle import oracle  # noqa: E402

OUT = ROOT / "tests" / "golden" / "workload_checksums.json"


def oracle_checksum(workload: str) -> dict:
    n, length, k, scaled, desc = WORKLOADS[workload]
    t0 = time.perf_counter()
    rows, cnts = [], []
    for g0 in range(0, n, 200):
        h, c = oracle.synth_sketch_batch(SEED, g0, min(200, n - g0), length, k, scaled)
        rows.append(h)
        cnts.append(c)
    counts = np.concatenate(cnts)
    cap = int(counts.max())
    hashes = np.concatenate([r[:, :cap] for r in rows])
    del rows
    t1 = time.perf_counter()
    ov = oracle.intersect_all(hashes, counts)  # int64 n x n, diagonal = sizes
    t2 = time.perf_counter()
    ov_ck = np.int64(0)
    with np.errstate(over="ignore"):
        for r0 in range(0, n, 1024):
            w = checksum_weights(np.arange(r0, min(n, r0 + 1024), dtype=np.int64), np.arange(n, dtype=np.int64))
            ov_ck = ov_ck + (ov[r0: r0 + 1024] * w).sum(dtype=np.int64)
        valid = np.arange(cap)[None, :] < counts[:, None]
        hash_ck = (hashes.view(np.int64) * valid).sum(dtype=np.int64)
    return {
        "workload": desc, "n_genomes": n, "genome_bp": length, "k": k, "scaled": scaled, "seed": SEED,
        "ov_weighted_sum": int(ov_ck), "hash_sum": int(hash_ck), "sketch_total": int(counts.sum()),
        "pairs_with_overlap": int((np.triu(ov, 1) > 0).sum()), "max_sketch": cap,
        "oracle_seconds": {"sketch": round(t1 - t0, 1), "intersect": round(t2 - t1, 1)},
        "threads": oracle.num_threads(),
    }


def main() -> None:
    names = sys.argv[1:] or ["config2", "config3"]
    data = json.loads(OUT.read_text()) if OUT.is_file() else {
        "_source": "tools/oracle_checksums.py: the CPU oracle over the COMPLETE workload (every genome, every pair)"}
    for name in names:
        data[name] = oracle_checksum(name)
        print(name, json.dumps(data[name]), flush=True)
        OUT.write_text(json.dumps(data, indent=1) + "\n")


if __name__ == "__main__":
    main()
