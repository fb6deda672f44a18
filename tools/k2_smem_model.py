#!/usr/bin/env python
"""Shared-memory wavefront model of the probing K2's inner loop (csrc/pairwise.cu, intersect_kernel).

No GPU needed.  Two random sketches of the configs[2] size are probed exactly as the kernel does it -- 32 lanes
take 32 consecutive hashes of the sorted subject column, look their bucket up in the 16-bit index (``idx``) and
compare the two staged query hashes ``seg[lo]``, ``seg[lo + 1]`` -- and every warp-wide load is priced with the
bank rules of the SM: 32 banks of 4 bytes; a 4-byte (or narrower) load costs as many wavefronts as the largest
number of distinct words any bank has to deliver; an 8-byte load is served per half-warp, 16 lanes at a time.

ncu on the real kernel (profiles/r02_k2probe_ncu.md): 629.6 M shared wavefronts for 78.4 M warp-probes = 8.03 per
warp-probe, 196.5 M of them (31 %) flagged as bank conflicts.  The model gives 7.5 (the staging of the segment
and the index build are not modelled), of which 5 are the minimum of the three loads.

Layouts compared:
  base      the kernel as built: idx u16 with ~3 buckets per element, seg u64
  sentinel  empty buckets point at the sentinel slot, so ~70 % of the lanes broadcast one address
  r2        2 buckets per element (idx span of a warp = 128 bytes instead of 192)
  planes    seg as two u32 planes: both low words, then the high word of a match only
"""
from __future__ import annotations

import sys

import numpy as np

MAXH = 18446744073709552  # scaled = 1000


def sketch(rng: np.random.Generator, n: int) -> np.ndarray:
    return np.unique(rng.integers(0, MAXH, size=n, dtype=np.uint64))


def make_plan(max_count: int, buckets_per_element: int = 3) -> tuple[int, int, int, int]:
    """seg_cap, R, pre, mul as make_plan() of pairwise.cu chooses them for one segment."""
    seg_cap = (max_count + 2 + 63) // 64 * 64
    r = buckets_per_element * seg_cap
    if buckets_per_element == 3:
        fit = (76288 - seg_cap * 8) // 2 - 2
        if fit < r and fit >= 2 * seg_cap:
            r = fit
    bl = int(MAXH).bit_length()
    pre = bl - 32 if bl > 32 else 0
    mul = min((r << 32) // ((MAXH >> pre) + 1), 0xFFFFFFFF)
    return seg_cap, r, pre, mul


def bucket(x: np.ndarray, pre: int, mul: int) -> np.ndarray:
    v = (x >> np.uint64(pre)) & np.uint64(0xFFFFFFFF)
    return ((v * np.uint64(mul)) >> np.uint64(32)).astype(np.int64)


def wavefronts(addr: np.ndarray, width: int, active: np.ndarray | None = None) -> int:
    """Wavefronts of one warp-wide shared-memory load of ``width`` bytes per lane at byte addresses ``addr``."""
    groups = [slice(0, 32)] if width <= 4 else [slice(0, 16), slice(16, 32)]
    total = 0
    for g in groups:
        a = addr[g] if active is None else addr[g][active[g]]
        if a.size == 0:
            continue
        words = np.unique(np.concatenate([(a + o) // 4 for o in range(0, max(width, 4), 4)]))
        total += int(np.bincount(words % 32, minlength=32).max())
    return total


def run(layout: str, n: int = 5026, warps: int = 3000, seed: int = 1) -> dict[str, float]:
    rng = np.random.default_rng(seed)
    q, s = sketch(rng, n), sketch(rng, n)
    # related genomes: a third of the subject's hashes are the query's (the match rate decides how often the
    # planes layout needs its third load)
    take = rng.choice(len(q), size=len(q) // 3, replace=False)
    s = np.unique(np.concatenate([s[: len(s) * 2 // 3], q[take]]))
    nq = len(q)
    seg_cap, r, pre, mul = make_plan(max(len(q), len(s)), 2 if layout == "r2" else 3)
    bq, bs = bucket(q, pre, mul), bucket(s, pre, mul)
    first_ge = np.searchsorted(bq, np.arange(r + 2), side="left")
    filled = np.bincount(bq, minlength=r + 2)[: r + 2]
    idx = np.where(filled > 0, first_ge, nq) if layout == "sentinel" else first_ge
    idx_base = seg_cap * 8
    cost = {"idx": 0, "seg0": 0, "seg1": 0, "verify": 0}
    starts = rng.integers(0, len(s) - 32, size=warps)
    for st in starts:
        b, x = bs[st: st + 32], s[st: st + 32]
        lo = idx[b]
        cost["idx"] += wavefronts(idx_base + 2 * b, 2)
        if layout == "planes":
            cost["seg0"] += wavefronts(4 * lo, 4)
            cost["seg1"] += wavefronts(4 * (lo + 1), 4)
            qpad = np.concatenate([q, np.full(2, np.iinfo(np.uint64).max, dtype=np.uint64)])
            m0 = (qpad[lo] & np.uint64(0xFFFFFFFF)) == (x & np.uint64(0xFFFFFFFF))
            m1 = (qpad[lo + 1] & np.uint64(0xFFFFFFFF)) == (x & np.uint64(0xFFFFFFFF))
            hit = m0 | m1
            cost["verify"] += wavefronts(4 * seg_cap + 4 * np.where(m0, lo, lo + 1), 4, hit)
        else:
            cost["seg0"] += wavefronts(8 * lo, 8)
            cost["seg1"] += wavefronts(8 * (lo + 1), 8)
    out = {k: v / warps for k, v in cost.items()}
    out["total"] = sum(out.values())
    out["crowded_buckets_pct"] = 100.0 * float((filled > 2).sum()) / max(1, int((filled > 0).sum()))
    return out


if __name__ == "__main__":
    for layout in sys.argv[1:] or ("base", "sentinel", "r2", "planes"):
        res = run(layout)
        print(f"{layout:9s}" + "  ".join(f"{k} {v:5.2f}" for k, v in res.items()))
