#!/bin/bash
# Multi-GPU check on a 2-GPU box:  gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_check2.sh final4'
# the 2-GPU CUDA tests (fused / NCCL exchange, sharded K2, `pyani-plus sourmash` under torchrun equal to the
# single-process database) and the bench line at N=2 launched as the driver launches it
set -u
OUT=gpurun_out
TAG=${1:-final}
mkdir -p $OUT
timeout 600 python -m pytest tests/test_multi_gpu_cuda.py -x -q > $OUT/${TAG}_tests_n2.log 2>&1; echo "rc=$?" >> $OUT/${TAG}_tests_n2.log
tail -3 $OUT/${TAG}_tests_n2.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 \
  bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/${TAG}_bench_n2.json 2> $OUT/${TAG}_bench_n2.err; echo "rc=$?" >> $OUT/${TAG}_bench_n2.err
python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_n2.json").read().strip().splitlines()[-1])
    e = d["e2e"]
    print("N=2 ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e ms", e.get("ms_per_step"), "stage", d["stage_ms"], "parity", d["parity"]["ok"])
except Exception as exc:
    print("no JSON:", exc)
PY
