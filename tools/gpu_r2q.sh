#!/bin/bash
set -u
OUT=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/r2q_tests.log 2>&1; echo "rc=$?" >> $OUT/r2q_tests.log
tail -3 $OUT/r2q_tests.log
bash tools/profile_r2.sh r02 config3
timeout 900 python bench.py > $OUT/r2q_bench_default.json 2> $OUT/r2q_bench_default.err; echo "rc=$?" >> $OUT/r2q_bench_default.err
for w in config2 config5; do timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > $OUT/r2q_bench_$w.json 2> $OUT/r2q_bench_$w.err; done
python - <<PY
import json
for v in ("default", "config2", "config5"):
    d = json.loads(open("$OUT/r2q_bench_%s.json" % v).read().strip().splitlines()[-1])
    print(v, round(d["value"]), round(d["ms_per_step"], 3), d["e2e"]["ms_per_step"], d["stage_ms"], d["roofline"]["frac"], d["roofline"]["traffic"], d["roofline"]["integer_pipes"], d["parity"]["ok"])
PY
