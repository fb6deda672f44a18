#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r2e_tests.log 2>&1; echo "rc=$?" >> $OUT/r2e_tests.log
for w in config3 config5 config4; do
  timeout 600 python tools/time_k1.py $w 5 > $OUT/r2e_time_$w.log 2>&1
done
tail -25 $OUT/r2e_tests.log; for w in config3 config5 config4; do cut -c1-700 $OUT/r2e_time_$w.log; done
