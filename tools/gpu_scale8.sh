#!/bin/bash
# the driver's scaling protocol on the default workload (configs[2]) + the north-star run (configs[3] on 8 GPUs)
set -u
OUT=gpurun_out
TAG=${1:-r02f}
mkdir -p $OUT
nvidia-smi -L | wc -l > $OUT/scale8_${TAG}_box.log; lscpu | grep -E "^CPU\(s\)|Model name|NUMA node\(s\)|Socket" >> $OUT/scale8_${TAG}_box.log; free -g | head -2 >> $OUT/scale8_${TAG}_box.log
nvidia-smi topo -m >> $OUT/scale8_${TAG}_box.log 2>&1
bash tools/gpu_scale.sh $TAG config3
run8() {  # name, extra args...
  local name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29577 \
    bench.py --gpus 8 --steps 6 --warmup 3 "$@" > $OUT/scale8_${TAG}_$name.json 2> $OUT/scale8_${TAG}_$name.err
  echo "$name rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("$OUT/scale8_${TAG}_$name.json").read().strip().splitlines()[-1])
    e = d["e2e"]
    print("  ms/step", round(d["ms_per_step"], 3), "e2e ms", e.get("ms_per_step"), e.get("ingest"), e.get("host_pack", {}).get("threads_per_rank"), "stage", d["stage_ms"], "parity", d["parity"]["ok"])
except Exception as exc:
    print("  no JSON:", exc)
PY
}
run8 c3_ht1 --workload config3 --host-threads 1
run8 c3_ht2 --workload config3 --host-threads 2
run8 c4 --workload config4
