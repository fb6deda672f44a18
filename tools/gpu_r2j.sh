#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/r2j_tests_all.log 2>&1; echo "rc=$?" >> $OUT/r2j_tests_all.log
tail -30 $OUT/r2j_tests_all.log
