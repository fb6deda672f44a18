#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
for v in default seed42 vote vote2 voteseed; do
  for w in config2 config3; do
    if [ $v = default ]; then
      timeout 300 python tools/time_k1.py $w 7 > $OUT/r2o_time_${v}_$w.log 2>&1
    else
      PANIB200_LIB=tools/variants/$v.so timeout 300 python tools/time_k1.py $w 7 > $OUT/r2o_time_${v}_$w.log 2>&1
    fi
  done
done
for f in $OUT/r2o_time_*.log; do echo $f; cut -c1-110 $f; done
