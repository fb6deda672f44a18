"""Host side of the ingest pipeline alone (no GPU): ASCII -> packed throughput of panib_pack_host_tiles at
several thread counts, next to a plain read of the same buffer (numpy sum), on THIS host.

    python tools/host_pack_bench.py [gigabases]
"""
import ctypes
import os
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pyani_plus_b200 import engine  # noqa: E402

lib = engine.load_library()
n = int(float(sys.argv[1]) * (1 << 30)) if len(sys.argv) > 1 else 1 << 30
n -= n % (64 * 4096)
rng = np.random.default_rng(1)
a = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n // 64)].repeat(64)
p = np.zeros(n // 16 + 16, np.uint32)
p = p[(-p.ctypes.data // 4) % 16:][: n // 16]  # 64-byte aligned: non-temporal stores
m = np.zeros(n // 32, np.uint32)
d = np.zeros(n // 4096, np.uint8)
cores = len(os.sched_getaffinity(0))
print(f"{n / 1e9:.2f} G bases, {cores} cores available, pool of {lib.panib_host_threads()} threads")
for th in sorted({1, 2, 4, 8, cores // 2, cores}):
    if th < 1 or th > cores:
        continue
    best = 1e9
    for _ in range(4):
        t = time.perf_counter()
        rc = lib.panib_pack_host_tiles(a.ctypes.data, n, p.ctypes.data, m.ctypes.data, d.ctypes.data, th)
        best = min(best, time.perf_counter() - t)
        assert rc == 0
    print(f"pack_host_tiles {th:3d} threads: {n / best / 1e9:7.2f} GB/s of ASCII ({best * 1e3:.1f} ms)")
