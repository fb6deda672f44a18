#!/bin/bash
set -u
OUT=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/r2w_tests.log 2>&1; echo "rc=$?" >> $OUT/r2w_tests.log
tail -3 $OUT/r2w_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 \
  bench.py --gpus 2 --steps 6 --warmup 3 > $OUT/r2w_bench_n2.json 2> $OUT/r2w_bench_n2.err; echo "rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2w_bench_n2.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("N=2 ms", round(d["ms_per_step"], 3), "e2e", round(e["ms_per_step"], 2), e["ingest"], "parity", d["parity"]["ok"])
PY
