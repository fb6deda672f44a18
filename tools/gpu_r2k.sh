#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
for v in warp; do
  PANIB200_LIB=tools/variants/$v.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > $OUT/r2k_tests_$v.log 2>&1; echo "rc=$?" >> $OUT/r2k_tests_$v.log
  for w in config2 config3 config5; do
    PANIB200_LIB=tools/variants/$v.so timeout 300 python tools/time_k1.py $w 5 > $OUT/r2k_time_${v}_$w.log 2>&1
  done
done
for w in config2 config3 config5; do timeout 300 python tools/time_k1.py $w 5 > $OUT/r2k_time_default_$w.log 2>&1; done
PANIB200_LIB=tools/variants/warp.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:sketch_hash_kernel -s 1 -c 1 \
    -f -o $OUT/prof_k1_r2k_warp python tools/time_k1.py config2 1 > $OUT/prof_k1_r2k_warp.log 2>&1
for f in $OUT/r2k_tests_*.log; do echo $f; tail -3 $f; done
for f in $OUT/r2k_time_*.log; do echo $f; cut -c1-120 $f; done
