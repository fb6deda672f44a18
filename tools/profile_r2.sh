#!/bin/bash
# Profiling recipe (B200_PROFILING.md) for the bench step, round 2.  Run under gpurun:
#   gpurun --timeout 1500 -- 'bash tools/profile_r2.sh r02 config3'
# Raw outputs land in gpurun_out/; then here:
#   python tools/ncu_to_json.py r02 config3 k1=gpurun_out/prof_k1_r02.ncu-rep k2_index=gpurun_out/prof_k2idx_r02.ncu-rep \
#          k2_probe=gpurun_out/prof_k2probe_r02.ncu-rep --sha gpurun_out/source_sha_r02.txt
# writes profiles/ncu_r02.json (what bench.py reads) and tools/summarise_ncu.py the text summaries.
set -u
TAG=${1:-r02}
WL=${2:-config3}
OUT=gpurun_out
mkdir -p $OUT
python -c "import bench, json; print(json.dumps({'all': bench.kernel_source_sha(), **{k: bench.kernel_source_sha(k) for k in bench.KERNEL_SOURCES}}))" > $OUT/source_sha_$TAG.txt
CMD="python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
EAGER="$CMD --no-graph"   # the full captures pick single launches: eager launches keep -s/-c counting simple
# 1) every launch of the bench command with its device time (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $OUT/launches_$TAG.csv $EAGER > $OUT/launches_$TAG.log 2>&1
# 2) full capture of K1 (one launch after warm-up) and of the K2 kernels of one step
ncu --set full --clock-control none --import-source on -k regex:sketch_hash_kernel -s 3 -c 1 \
    -f -o $OUT/prof_k1_$TAG $EAGER > $OUT/prof_k1_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"index_|DeviceRadixSort|DeviceScan|sort_" -s 40 -c 20 \
    -f -o $OUT/prof_k2idx_$TAG $EAGER --k2 index > $OUT/prof_k2idx_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:intersect_kernel -s 3 -c 1 \
    -f -o $OUT/prof_k2probe_$TAG $EAGER --k2 probe > $OUT/prof_k2probe_$TAG.log 2>&1
ls -la $OUT | grep $TAG
