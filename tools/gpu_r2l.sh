#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > $OUT/r2l_tests.log 2>&1; echo "rc=$?" >> $OUT/r2l_tests.log
for w in config2 config3 config5; do timeout 300 python tools/time_k1.py $w 7 > $OUT/r2l_time_$w.log 2>&1; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sketch_hash_kernel -s 1 -c 1 \
    -f -o $OUT/prof_k1_r2l python tools/time_k1.py config2 1 > $OUT/prof_k1_r2l.log 2>&1
tail -3 $OUT/r2l_tests.log
for f in $OUT/r2l_time_*.log; do echo $f; cut -c1-120 $f; done
