#!/bin/bash
set -u
OUT=gpurun_out
TAG=final2
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_tests.log 2>&1; echo "rc=$?" >> $OUT/${TAG}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "rc=$?" >> $OUT/${TAG}_smoke.log
bash tools/profile_r2.sh r02 config3 > /dev/null 2>&1
timeout 900 python bench.py > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err; echo "rc=$?" >> $OUT/${TAG}_bench_default.err
timeout 900 python bench.py --impl reference > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; echo "rc=$?" >> $OUT/${TAG}_bench_reference.err
tail -3 $OUT/${TAG}_tests.log; tail -2 $OUT/${TAG}_smoke.log
python - <<PY
import json
for v in ("default", "reference"):
    try:
        d = json.loads(open("$OUT/${TAG}_bench_%s.json" % v).read().strip().splitlines()[-1])
        e = d.get("e2e") or {}
        print(v, "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e ms", e.get("ms_per_step"), e.get("ingest"), "roof", (d.get("roofline") or {}).get("frac"), (d.get("roofline") or {}).get("traffic"), "parity", (d.get("parity") or {}).get("ok"))
    except Exception as exc:
        print(v, "failed", exc)
PY
