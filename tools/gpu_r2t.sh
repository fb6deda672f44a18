#!/bin/bash
set -u
OUT=gpurun_out
hostname > $OUT/r2t_box.log; lscpu | grep -E "^CPU\(s\)|MHz|L3|Model name" >> $OUT/r2t_box.log; uptime >> $OUT/r2t_box.log
timeout 300 python tools/host_pack_bench.py 2 >> $OUT/r2t_box.log 2>&1
B="python bench.py --workload config3 --steps 6 --warmup 3 --no-cpu-baseline"
for rep in 1 2; do
for mb in 0 16; do
  PANIB_INGEST_RING_MB=$mb timeout 600 $B > $OUT/r2t_bench_ring${mb}_$rep.json 2> $OUT/r2t_bench_ring${mb}_$rep.err
  PANIB_INGEST_RING_MB=$mb PANIB_INGEST_RAW=0 timeout 600 $B > $OUT/r2t_bench_ring${mb}noraw_$rep.json 2> $OUT/r2t_bench_ring${mb}noraw_$rep.err
done
done
cat $OUT/r2t_box.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2t_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d["e2e"]
        print(f.split("bench_")[1], "e2e ms", round(e["ms_per_step"], 2), {k: e["ingest"][k] for k in ("h2d_bytes", "chunks_as_ascii")})
    except Exception as exc:
        print(f, "failed", exc)
PY
