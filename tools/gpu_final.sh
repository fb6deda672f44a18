#!/bin/bash
# the round-end protocol, run by hand: GPU tests, smoke, both bench arms on the default workload, the other
# BASELINE configs on one GPU, and the drop-in CLI wall times
set -u
OUT=gpurun_out
TAG=${1:-final}
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_tests.log 2>&1; echo "rc=$?" >> $OUT/${TAG}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "rc=$?" >> $OUT/${TAG}_smoke.log
timeout 900 python bench.py > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err; echo "rc=$?" >> $OUT/${TAG}_bench_default.err
timeout 900 python bench.py --impl reference > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; echo "rc=$?" >> $OUT/${TAG}_bench_reference.err
for w in config2 config5 config4; do
  timeout 900 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_$w.json 2> $OUT/${TAG}_bench_$w.err; echo "rc=$?" >> $OUT/${TAG}_bench_$w.err
done
timeout 900 python tools/cli_walltime.py 1000 5000000 1000 > $OUT/${TAG}_cli_1000.log 2>&1
timeout 1500 python tools/cli_walltime.py 10000 500000 100 cold > $OUT/${TAG}_cli_10000.log 2>&1; echo "rc=$?" >> $OUT/${TAG}_cli_10000.log
tail -3 $OUT/${TAG}_tests.log; tail -2 $OUT/${TAG}_smoke.log
python - <<PY
import json
for v in ("default", "reference", "config2", "config5", "config4"):
    try:
        d = json.loads(open("$OUT/${TAG}_bench_%s.json" % v).read().strip().splitlines()[-1])
        e = d.get("e2e") or {}
        print(v, "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e ms", e.get("ms_per_step"), "stage", d.get("stage_ms"), "roof", (d.get("roofline") or {}).get("frac"), (d.get("roofline") or {}).get("traffic"), "parity", (d.get("parity") or {}).get("ok"))
    except Exception as exc:
        print(v, "failed", exc)
PY
grep -v "━" $OUT/${TAG}_cli_1000.log | tail -4; grep -v "━" $OUT/${TAG}_cli_10000.log | tail -4
