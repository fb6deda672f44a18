#!/bin/bash
set -u
OUT=gpurun_out
B="python bench.py --workload config3 --steps 6 --warmup 3 --no-cpu-baseline"
for rep in 1 2; do
  PANIB_INGEST_RING_MB=0 timeout 600 $B > $OUT/r2v_bench_ring0_$rep.json 2> $OUT/r2v_bench_ring0_$rep.err
  PANIB_INGEST_RING_MB=16 timeout 600 $B > $OUT/r2v_bench_ring16_$rep.json 2> $OUT/r2v_bench_ring16_$rep.err
  PANIB_INGEST_RING_MB=16 PANIB_INGEST_CLWB=1 timeout 600 $B > $OUT/r2v_bench_ring16clwb_$rep.json 2> $OUT/r2v_bench_ring16clwb_$rep.err
  PANIB_INGEST_RING_MB=32 PANIB_INGEST_CLWB=1 timeout 600 $B > $OUT/r2v_bench_ring32clwb_$rep.json 2> $OUT/r2v_bench_ring32clwb_$rep.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2v_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d["e2e"]
        print(f.split("bench_")[1], "e2e ms", round(e["ms_per_step"], 2), {k: e["ingest"][k] for k in ("h2d_bytes", "chunks_as_ascii", "ranks_using_the_cached_ring")})
    except Exception as exc:
        print(f, "failed", exc)
PY
