#!/bin/bash
# scaling of the default bench (configs[2]) over the GPUs of this box: N = 1, 2, 4, 8 as far as there are GPUs
set -u
OUT=gpurun_out
TAG=${1:-r02}
WL=${2:-config3}
mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l)
for N in 1 2 4 8; do
  [ $N -gt $NG ] && break
  if [ $N -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --workload $WL --steps 10 --warmup 3 > $OUT/scale_${TAG}_${WL}_n$N.json 2> $OUT/scale_${TAG}_${WL}_n$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$N \
      bench.py --gpus $N --workload $WL --steps 10 --warmup 3 > $OUT/scale_${TAG}_${WL}_n$N.json 2> $OUT/scale_${TAG}_${WL}_n$N.err
  fi
  echo "N=$N rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("$OUT/scale_${TAG}_${WL}_n$N.json").read().strip().splitlines()[-1])
    print("  ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e ms", d["e2e"].get("ms_per_step"), "stage", d["stage_ms"], "parity", d["parity"]["ok"], "exchange", d.get("exchange"))
except Exception as e:
    print("  no JSON:", e)
PY
done
