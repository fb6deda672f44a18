#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python tools/cli_walltime.py 10000 500000 100 cold > $OUT/r2n_cli_10000.log 2>&1; echo "rc=$?" >> $OUT/r2n_cli_10000.log
tail -6 $OUT/r2n_cli_10000.log
bash tools/sanitize.sh
