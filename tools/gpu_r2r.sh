#!/bin/bash
set -u
OUT=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "ingest or pipeline" > $OUT/r2r_tests.log 2>&1; echo "rc=$?" >> $OUT/r2r_tests.log
tail -4 $OUT/r2r_tests.log
B="python bench.py --workload config3 --steps 6 --warmup 3 --no-cpu-baseline"
for mb in 16 0 4 8 32 64; do
  PANIB_INGEST_RING_MB=$mb timeout 600 $B > $OUT/r2r_bench_ring$mb.json 2> $OUT/r2r_bench_ring$mb.err
done
PANIB_INGEST_RING_MB=16 PANIB_INGEST_RAW=0 timeout 600 $B > $OUT/r2r_bench_ring16noraw.json 2> $OUT/r2r_bench_ring16noraw.err
python - <<'PY'
import json
for v in ("ring16", "ring0", "ring4", "ring8", "ring32", "ring64", "ring16noraw"):
    try:
        d = json.loads(open(f"gpurun_out/r2r_bench_{v}.json").read().strip().splitlines()[-1])
        e = d["e2e"]
        print(v, "value ms", round(d["ms_per_step"], 2), "e2e ms", round(e["ms_per_step"], 2), {k: e["ingest"][k] for k in ("h2d_bytes", "chunks_as_ascii", "dirty_tiles")}, "parity", d["parity"]["ok"])
    except Exception as exc:
        print(v, "failed", exc)
PY
