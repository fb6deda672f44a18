#!/bin/bash
# round-2 first GPU pass: parity suite, K1 A/B against the round-1 library, one full ncu capture of K1
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r2a_tests.log 2>&1
echo "pytest rc=$?" >> $OUT/r2a_tests.log
for w in config2 config5; do
  timeout 300 python tools/time_k1.py $w 7 > $OUT/r2a_time_new_$w.log 2>&1
  PANIB200_LIB=tools/variants/r1.so timeout 300 python tools/time_k1.py $w 7 > $OUT/r2a_time_r1_$w.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sketch_hash_kernel -s 1 -c 1 \
    -f -o $OUT/prof_k1_r2a python tools/time_k1.py config2 1 > $OUT/prof_k1_r2a.log 2>&1
tail -3 $OUT/r2a_tests.log; cat $OUT/r2a_time_*.log
