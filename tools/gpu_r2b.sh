#!/bin/bash
# K1 variant A/B: parity subset + timing per variant, ncu of the default
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > $OUT/r2b_tests_default.log 2>&1; echo "rc=$?" >> $OUT/r2b_tests_default.log
timeout 300 python tools/time_k1.py config2 7 > $OUT/r2b_time_default_config2.log 2>&1
timeout 300 python tools/time_k1.py config5 3 > $OUT/r2b_time_default_config5.log 2>&1
for v in t128 t128m7 m3 r1; do
  PANIB200_LIB=tools/variants/$v.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > $OUT/r2b_tests_$v.log 2>&1; echo "rc=$?" >> $OUT/r2b_tests_$v.log
  PANIB200_LIB=tools/variants/$v.so timeout 300 python tools/time_k1.py config2 7 > $OUT/r2b_time_${v}_config2.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sketch_hash_kernel -s 1 -c 1 \
    -f -o $OUT/prof_k1_r2b python tools/time_k1.py config2 1 > $OUT/prof_k1_r2b.log 2>&1
PANIB200_LIB=tools/variants/t128.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:sketch_hash_kernel -s 1 -c 1 \
    -f -o $OUT/prof_k1_r2b_t128 python tools/time_k1.py config2 1 > $OUT/prof_k1_r2b_t128.log 2>&1
for f in $OUT/r2b_tests_*.log; do echo $f; tail -2 $f; done
for f in $OUT/r2b_time_*.log; do echo $f; cut -c1-120 $f; done
