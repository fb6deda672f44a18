#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
lscpu | grep -E "Model name|Socket|Core|Thread|NUMA" > $OUT/r2c_host.txt; nproc >> $OUT/r2c_host.txt; free -g >> $OUT/r2c_host.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r2c_tests.log 2>&1; echo "rc=$?" >> $OUT/r2c_tests.log
python - > $OUT/r2c_hostpack.log 2>&1 <<'PY'
import ctypes, numpy as np, time, torch
from pyani_plus_b200 import engine
L = engine.load_library()
n = 500_000_000 // 32 * 32
a = np.frombuffer(b"ACGT", dtype=np.uint8)[np.random.default_rng(2).integers(0, 4, n, dtype=np.uint8)]
hp = torch.empty(n // 16, dtype=torch.int32, pin_memory=True); hm = torch.empty(n // 32, dtype=torch.int32, pin_memory=True)
for t in (1, 2, 4, 8, 16, 32, 0):
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter(); L.panib_pack_host(a.ctypes.data, n, hp.data_ptr(), hm.data_ptr(), t); best = min(best, time.perf_counter() - t0)
    print(f"pack_host threads={t}: {best*1e3:.1f} ms = {n/best/1e9:.1f} GB/s of ASCII", flush=True)
print("pool", L.panib_host_threads())
PY
timeout 900 python bench.py --workload config3 --steps 10 --warmup 3 > $OUT/r2c_bench_config3.json 2> $OUT/r2c_bench_config3.err; echo "rc=$?" >> $OUT/r2c_bench_config3.err
timeout 600 python bench.py --workload config2 --steps 10 --warmup 3 > $OUT/r2c_bench_config2.json 2> $OUT/r2c_bench_config2.err; echo "rc=$?" >> $OUT/r2c_bench_config2.err
timeout 900 python bench.py --impl reference --workload config3 --steps 3 --warmup 1 > $OUT/r2c_ref_config3.json 2> $OUT/r2c_ref_config3.err
tail -3 $OUT/r2c_tests.log; cat $OUT/r2c_hostpack.log $OUT/r2c_host.txt; tail -2 $OUT/r2c_bench_config3.err
