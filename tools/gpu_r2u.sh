#!/bin/bash
set -u
OUT=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "ingest or pipeline" > $OUT/r2u_tests.log 2>&1; echo "rc=$?" >> $OUT/r2u_tests.log
tail -3 $OUT/r2u_tests.log
B="python bench.py --workload config3 --steps 6 --warmup 3 --no-cpu-baseline"
for rep in 1 2; do
  timeout 600 $B > $OUT/r2u_bench_auto_$rep.json 2> $OUT/r2u_bench_auto_$rep.err
  PANIB_INGEST_RING_MB=0 timeout 600 $B > $OUT/r2u_bench_ring0_$rep.json 2> $OUT/r2u_bench_ring0_$rep.err
  PANIB_INGEST_RING_MB=16 timeout 600 $B > $OUT/r2u_bench_ring16_$rep.json 2> $OUT/r2u_bench_ring16_$rep.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2u_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d["e2e"]
        print(f.split("bench_")[1], "e2e ms", round(e["ms_per_step"], 2), {k: e["ingest"][k] for k in ("h2d_bytes", "chunks_as_ascii", "ranks_using_the_cached_ring")})
    except Exception as exc:
        print(f, "failed", exc)
PY
