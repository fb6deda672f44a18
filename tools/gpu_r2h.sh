#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/r2h_box.log; lscpu | head -25 >> $OUT/r2h_box.log; numactl -H >> $OUT/r2h_box.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/r2h_tests.log 2>&1; echo "rc=$?" >> $OUT/r2h_tests.log
for w in config2 config3 config5; do
  timeout 300 python tools/time_k1.py $w 5 > $OUT/r2h_time_$w.log 2>&1
  PANIB_NO_WORKSPACE=1 timeout 300 python tools/time_k1.py $w 5 > $OUT/r2h_time_nows_$w.log 2>&1
done
timeout 600 python bench.py --workload config3 --steps 10 --warmup 3 > $OUT/r2h_bench_config3.json 2> $OUT/r2h_bench_config3.err; echo "rc=$?" >> $OUT/r2h_bench_config3.err
timeout 600 python bench.py --impl reference --workload config3 --steps 2 --warmup 1 > $OUT/r2h_ref_config3.json 2> $OUT/r2h_ref_config3.err; echo "rc=$?" >> $OUT/r2h_ref_config3.err
tail -5 $OUT/r2h_tests.log; for w in config2 config3 config5; do cut -c1-200 $OUT/r2h_time_$w.log; cut -c1-200 $OUT/r2h_time_nows_$w.log; done; tail -2 $OUT/r2h_bench_config3.err; cut -c1-400 $OUT/r2h_bench_config3.json
