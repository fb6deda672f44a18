#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > $OUT/r2g_tests.log 2>&1; echo "rc=$?" >> $OUT/r2g_tests.log
timeout 300 python tools/time_k1.py config2 7 > $OUT/r2g_time_default_config2.log 2>&1
timeout 300 python tools/time_k1.py config3 5 > $OUT/r2g_time_default_config3.log 2>&1
timeout 300 python tools/time_k1.py config4 3 > $OUT/r2g_time_default_config4.log 2>&1
for v in g2 g2m3; do
  PANIB200_LIB=tools/variants/$v.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fixture or kmer or config2" > $OUT/r2g_tests_$v.log 2>&1; echo "rc=$?" >> $OUT/r2g_tests_$v.log
  PANIB200_LIB=tools/variants/$v.so timeout 300 python tools/time_k1.py config2 7 > $OUT/r2g_time_${v}_config2.log 2>&1
done
PANIB200_LIB=tools/variants/g2.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:sketch_hash_kernel -s 1 -c 1 \
    -f -o $OUT/prof_k1_r2g_g2 python tools/time_k1.py config2 1 > $OUT/prof_k1_r2g_g2.log 2>&1
timeout 900 python tools/cli_walltime.py 1000 5000000 1000 > $OUT/r2g_cli_1000.log 2>&1
for f in $OUT/r2g_tests*.log; do echo $f; tail -2 $f; done
for f in $OUT/r2g_time_*.log; do echo $f; cut -c1-420 $f; done
cat $OUT/r2g_cli_1000.log | tail -8
