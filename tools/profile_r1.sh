#!/bin/bash
# Profiling recipe (B200_PROFILING.md) for the bench step.  Run under gpurun:
#   gpurun --timeout 1500 -- 'bash tools/profile_r1.sh r01'
# Raw outputs land in gpurun_out/; `python tools/summarise_ncu.py r01` writes the summaries kept in profiles/.
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
EAGER="$CMD --no-graph"   # the full captures pick single launches: eager launches keep -s/-c counting simple
# 1) every launch of the default bench command with its device time (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $OUT/launches_$TAG.csv $CMD > $OUT/launches_$TAG.log 2>&1
if [ "$(grep -c sketch_hash_kernel $OUT/launches_$TAG.csv)" -lt 4 ]; then  # graph replay not visible to ncu: eager list
    ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
        --log-file $OUT/launches_$TAG.csv $EAGER > $OUT/launches_$TAG.log 2>&1
fi
# 2) full capture of the two hot kernels at config 2 (one launch each, after warm-up)
ncu --set full --clock-control none --import-source on -k regex:sketch_hash_kernel -s 3 -c 1 \
    -f -o $OUT/prof_k1_$TAG $EAGER > $OUT/prof_k1_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:intersect_kernel -s 3 -c 1 \
    -f -o $OUT/prof_k2_$TAG $EAGER > $OUT/prof_k2_$TAG.log 2>&1
# 3) K2 where it matters: 1,000 genomes (config 3), 499,500 pairs in one launch
ncu --set full --clock-control none --import-source on -k regex:intersect_kernel -s 1 -c 1 \
    -f -o $OUT/prof_k2c3_$TAG python tools/time_k1.py config3 1 > $OUT/prof_k2c3_$TAG.log 2>&1
# 4) K2, inverted-index form at config 3: every launch with its time (sort / scan / index_* kernels), and a
#    full capture of the AND+POPC bit-matrix kernel
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/launches_idx_$TAG.csv python tools/time_k1.py config3 1 > $OUT/launches_idx_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:index_dense_kernel -s 1 -c 1 \
    -f -o $OUT/prof_k2idx_$TAG python tools/time_k1.py config3 1 > $OUT/prof_k2idx_$TAG.log 2>&1
ls -la $OUT | grep $TAG
