#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > $OUT/r2f_tests.log 2>&1; echo "rc=$?" >> $OUT/r2f_tests.log
timeout 300 python tools/time_k1.py config2 7 > $OUT/r2f_time_config2.log 2>&1
timeout 300 python tools/time_k1.py config5 3 > $OUT/r2f_time_config5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sketch_hash_kernel -s 1 -c 1 \
    -f -o $OUT/prof_k1_r2f python tools/time_k1.py config2 1 > $OUT/prof_k1_r2f.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/launches_idx_r2f.csv python tools/time_k1.py config3 1 > $OUT/launches_idx_r2f.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:index_ -c 60 --csv \
    --log-file $OUT/launches_idx_c4_r2f.csv python tools/time_k1.py config4 1 > $OUT/launches_idx_c4_r2f.log 2>&1
tail -3 $OUT/r2f_tests.log; cut -c1-300 $OUT/r2f_time_config2.log $OUT/r2f_time_config5.log
