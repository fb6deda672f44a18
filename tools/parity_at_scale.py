#!/usr/bin/env python
"""Parity at BASELINE sizes (SURVEY.md 8d "Parity at scale"): CUDA path vs the CPU oracle.

    python tools/parity_at_scale.py --workload config3            # every sketch, the full count matrix
    python tools/parity_at_scale.py --workload config5
    python tools/parity_at_scale.py --workload config4 --sample 128   # sampled genomes + checksums

The GPU sketches all N synthetic genomes and intersects all pairs.  The oracle (C port, all host
threads) regenerates the checked genomes from the same counter-based recipe, sketches them and
intersects them; hashes and counts must be bit-exact, ANI within 1e-12 (CUDA pow vs host pow).
For a sample, the checked genomes are blocks of 16 ids spread evenly over the id range and the sampled sub-matrix of
the FULL GPU matrix is compared; size-independent properties are checked on the whole matrix
(symmetry, diagonal = sketch size, ov <= min size).  One JSON line on stdout; test infrastructure,
not product code.
"""

from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

from bench import SEED, WORKLOADS  # noqa: E402
from oracle import oracle  # noqa: E402


def check_workload(workload: str, sample: int = 0, batch: int = 250, k2_method: str = "auto") -> dict:
    """Run one BASELINE workload on cuda:0, check it against the oracle, return the report dict."""
    import torch

    from pyani_plus_b200 import engine

    n, length, k, scaled, desc = WORKLOADS[workload]
    eng = engine.Engine(0)
    t0 = time.perf_counter()
    parts = []
    for g0 in range(0, n, batch):
        m = min(batch, n - g0)
        d_ascii, tile_off = eng.synth_ascii_stream(SEED, g0, m, length)
        parts.append(eng.sketch_ascii_stream(d_ascii, tile_off, k, scaled, from_host=False))
        del d_ascii
    table = parts[0] if len(parts) == 1 else eng.concat_tables(parts, k, scaled)
    del parts
    ov_d = eng.intersect(table, method=k2_method)
    k2_used = eng.last_intersect_method
    ident_d, cov_d = eng.ani_device(ov_d, table)
    torch.cuda.synchronize()
    gpu_s = time.perf_counter() - t0
    ov = ov_d.cpu().numpy().astype(np.int64)
    counts = table.counts.cpu().numpy().astype(np.int64)

    # ---- whole-matrix properties (size independent)
    props = {
        "symmetric": bool((ov == ov.T).all()),
        "diag_is_size": bool((np.diag(ov) == counts).all()),
        "ov_le_min_size": bool((ov <= np.minimum.outer(counts, counts)).all()),
        "null_pairs": int((ov == 0).sum()),
        "matrix_sum": int(ov.sum()),
    }

    # ---- oracle on the checked genomes
    if not sample or sample >= n:
        ids = np.arange(n)
    else:  # blocks of 16 consecutive ids (one oracle batch call each, all host threads busy) spread over 0..n
        blk = 16
        starts = np.linspace(0, n - blk, max(1, sample // blk)).round().astype(np.int64)
        ids = np.unique((starts[:, None] + np.arange(blk)[None, :]).ravel())
    t1 = time.perf_counter()
    runs = np.split(ids, np.where(np.diff(ids) != 1)[0] + 1)  # consecutive id runs -> one batch call each
    rows, cnts = [], []
    for run in runs:
        for b0 in range(0, len(run), 100):
            h, c = oracle.synth_sketch_batch(SEED, int(run[b0]), len(run[b0: b0 + 100]), length, k, scaled)
            rows.append(h)
            cnts.append(c)
    cap = max(h.shape[1] for h in rows)
    want_h = np.zeros((len(ids), cap), dtype=np.uint64)
    r0 = 0
    for h in rows:
        want_h[r0: r0 + h.shape[0], : h.shape[1]] = h
        r0 += h.shape[0]
    want_c = np.concatenate(cnts).astype(np.int64)
    sketch_s = time.perf_counter() - t1

    got_rows = table.rows[torch.from_numpy(ids).to(table.rows.device)].cpu().numpy().view(np.uint64)
    hash_mismatch = 0
    checked_hashes = 0
    for i, g in enumerate(ids):
        a, b = got_rows[i, : counts[g]], want_h[i, : want_c[i]]
        checked_hashes += len(b)
        if len(a) != len(b) or not (a == b).all():
            hash_mismatch += 1
    t2 = time.perf_counter()
    want_ov = oracle.intersect_all(want_h, want_c)
    inter_s = time.perf_counter() - t2
    sub = ov[np.ix_(ids, ids)]
    count_mismatch = int((sub != want_ov).sum())

    ident_h, cov_h = engine.ani_host(sub.astype(np.uint32), counts[ids], counts[ids], k)
    idt = torch.from_numpy(ids).to(ident_d.device)
    ident_s = ident_d[idt][:, idt].cpu().numpy()
    cov_s = cov_d[idt][:, idt].cpu().numpy()
    nan_equal = bool((np.isnan(ident_s) == np.isnan(ident_h)).all() and (np.isnan(cov_s) == np.isnan(cov_h)).all())
    ani_err = float(max(np.nanmax(np.abs(ident_s - ident_h), initial=0.0), np.nanmax(np.abs(cov_s - cov_h), initial=0.0)))
    # the host formula against the oracle's own pair rows on a few hundred pairs
    rng = np.random.default_rng(1)
    row_mismatch = 0
    for _ in range(300):
        i, j = (int(x) for x in rng.integers(0, len(ids), 2))
        row = oracle.pair_row(int(want_ov[i, j]), int(want_c[i]), int(want_c[j]), k)
        if row is None:
            row_mismatch += 0 if np.isnan(ident_h[i, j]) and np.isnan(cov_h[i, j]) else 1
        else:
            row_mismatch += 0 if (ident_h[i, j] == row["max_containment_ani"]
                                  and cov_h[i, j] == row["query_containment_ani"]) else 1

    # the COMPLETE result against the oracle's pinned checksum of the complete workload (every hash, every pair)
    from bench import checksum_weights, expected_checksum  # noqa: PLC0415

    with np.errstate(over="ignore"):
        ov_ck = np.int64(0)
        cols = np.arange(n, dtype=np.int64)
        for r0 in range(0, n, 1024):
            w = checksum_weights(np.arange(r0, min(n, r0 + 1024), dtype=np.int64), cols)
            ov_ck = ov_ck + (ov[r0: r0 + 1024] * w).sum(dtype=np.int64)
    valid = torch.arange(table.rows.shape[1], device=table.rows.device)[None, :] < table.counts[:, None]
    checksum = {"ov_weighted_sum": int(ov_ck), "hash_sum": int((table.rows * valid).sum().item()),
                "sketch_total": int(counts.sum())}
    pinned = expected_checksum(workload)
    checksum_ok = pinned is None or checksum == pinned

    ok = (hash_mismatch == 0 and count_mismatch == 0 and nan_equal and ani_err <= 1e-12 and row_mismatch == 0
          and props["symmetric"] and props["diag_is_size"] and props["ov_le_min_size"] and checksum_ok)
    return {
        "workload": desc, "n_genomes": n, "genome_bp": length, "k": k, "scaled": scaled, "seed": SEED,
        "checked_genomes": len(ids), "checked_hashes": checked_hashes, "checked_pairs_ordered": int(sub.size),
        "sketches_with_mismatch": hash_mismatch, "count_cells_with_mismatch": count_mismatch,
        "ani_null_pattern_equal": nan_equal, "ani_device_vs_host_max_abs_err": ani_err,
        "ani_host_rows_not_equal_oracle": row_mismatch, "whole_matrix": props,
        "result_checksum": checksum, "oracle_checksum_of_complete_workload": pinned, "checksum_ok": checksum_ok,
        "sketch_sizes": {"min": int(counts.min()), "mean": float(counts.mean()), "max": int(counts.max())},
        "seconds": {"gpu_generate_sketch_intersect": gpu_s, "oracle_sketch": sketch_s, "oracle_intersect": inter_s},
        "k2_method": k2_used, "oracle_threads": oracle.num_threads(), "library": engine.library_version(), "ok": ok}


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--workload", default="config3", choices=sorted(WORKLOADS))
    ap.add_argument("--sample", type=int, default=0, help="genomes checked by the oracle (0 = all)")
    ap.add_argument("--batch", type=int, default=250, help="genomes generated + sketched per GPU pass")
    ap.add_argument("--k2", default="auto", choices=["auto", "probe", "index"])
    args = ap.parse_args()
    report = check_workload(args.workload, args.sample, args.batch, args.k2)
    print(json.dumps(report), flush=True)
    sys.exit(0 if report["ok"] else 1)


if __name__ == "__main__":
    main()
