#!/bin/bash
set -u
OUT=gpurun_out
for v in vote2seed seed42; do
  for w in config3 config5; do
    PANIB200_LIB=tools/variants/$v.so timeout 300 python tools/time_k1.py $w 7 > $OUT/r2p_time_${v}_$w.log 2>&1
  done
done
timeout 300 python tools/time_k1.py config5 7 > $OUT/r2p_time_default_config5.log 2>&1
for f in $OUT/r2p_time_*.log; do echo $f; cut -c1-110 $f; done
