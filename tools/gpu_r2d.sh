#!/bin/bash
# 2-GPU pass: full GPU suite (incl. the multi-GPU tests), 2-GPU bench at config3
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/r2d_tests.log 2>&1; echo "rc=$?" >> $OUT/r2d_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 \
   bench.py --gpus 2 --workload config3 --steps 10 --warmup 3 > $OUT/r2d_bench_config3_n2.json 2> $OUT/r2d_bench_config3_n2.err; echo "rc=$?" >> $OUT/r2d_bench_config3_n2.err
tail -30 $OUT/r2d_tests.log; tail -5 $OUT/r2d_bench_config3_n2.err; cut -c1-600 $OUT/r2d_bench_config3_n2.json
