#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python tools/host_pack_bench.py 2 > $OUT/r2i_hostpack.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "ingest or pipeline" > $OUT/r2i_tests.log 2>&1; echo "rc=$?" >> $OUT/r2i_tests.log
B="python bench.py --workload config3 --steps 6 --warmup 3 --no-cpu-baseline"
timeout 600 $B > $OUT/r2i_bench_default.json 2> $OUT/r2i_bench_default.err
PANIB_INGEST_RAW=0 timeout 600 $B > $OUT/r2i_bench_noraw.json 2> $OUT/r2i_bench_noraw.err
PANIB_INGEST_SPARSE=0 timeout 600 $B > $OUT/r2i_bench_dense.json 2> $OUT/r2i_bench_dense.err
PANIB_INGEST_RAW=2 timeout 600 $B > $OUT/r2i_bench_raweager.json 2> $OUT/r2i_bench_raweager.err
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r2i_tests_all.log 2>&1; echo "rc=$?" >> $OUT/r2i_tests_all.log
cat $OUT/r2i_hostpack.log; tail -4 $OUT/r2i_tests.log; tail -3 $OUT/r2i_tests_all.log
python - <<'PY'
import json
for v in ("default", "noraw", "dense", "raweager"):
    try:
        d = json.loads(open(f"gpurun_out/r2i_bench_{v}.json").read().strip().splitlines()[-1])
        e = d["e2e"]
        print(v, "value ms", round(d["ms_per_step"], 2), "e2e ms", round(e["ms_per_step"], 2), e.get("ingest"), "parity", d["parity"]["ok"])
    except Exception as exc:
        print(v, "failed", exc)
PY
