#!/bin/bash
# Profiling recipe (B200_PROFILING.md) for the config-2 bench step.  Run under gpurun:
#   gpurun --timeout 1500 -- 'bash tools/profile_r1.sh r01'
# Outputs land in gpurun_out/; summaries are copied to profiles/ by tools/summarise_ncu.py.
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
# 1) every launch with its device time (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/launches_$TAG.csv $CMD > $OUT/launches_$TAG.log 2>&1
# 2) full capture of the two hot kernels (one launch each, after warm-up)
ncu --set full --clock-control none --import-source on -k regex:sketch_hash_kernel -s 3 -c 1 \
    -f -o $OUT/prof_k1_$TAG $CMD > $OUT/prof_k1_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:intersect_kernel -s 3 -c 1 \
    -f -o $OUT/prof_k2_$TAG $CMD > $OUT/prof_k2_$TAG.log 2>&1
ls -la $OUT
