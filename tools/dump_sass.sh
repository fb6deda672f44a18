#!/bin/bash
# SASS listings of the hot kernels of the built library -> profiles/<tag>/sass_*.txt (no GPU needed).
#   bash tools/dump_sass.sh r02
TAG=${1:-r02}
LIB=pyani_plus_b200/libpanib200.so
mkdir -p profiles/$TAG
for k in sketch_hash_kernelILi31ELb1 intersect_kernelILb0 index_insert_kernel index_classify_kernel index_dense_kernel \
         index_sparse_kernel sketch_compact_scatter_kernel sketch_sort_buckets_kernel; do
  sym=$(cuobjdump -sass $LIB | grep "Function :" | grep "$k" | head -1 | awk '{print $3}')
  [ -z "$sym" ] && continue
  cuobjdump -sass -fun "$sym" $LIB 2>/dev/null | grep -v "^$" > profiles/$TAG/sass_$k.txt
  n=$(grep -cE "^\s+/\*[0-9a-f]{4,5}\*/" profiles/$TAG/sass_$k.txt)
  echo "$k: $n instructions; LDGSTS $(grep -c LDGSTS profiles/$TAG/sass_$k.txt), LDS.64 $(grep -c 'LDS.64' profiles/$TAG/sass_$k.txt), STS.64 $(grep -c 'STS.64' profiles/$TAG/sass_$k.txt), IMAD.WIDE $(grep -c 'IMAD.WIDE' profiles/$TAG/sass_$k.txt), POPC $(grep -c POPC profiles/$TAG/sass_$k.txt), ATOM $(grep -c 'ATOM\|RED' profiles/$TAG/sass_$k.txt), UTMALDG/UTCMMA $(grep -c 'UTMALDG\|UTC.MMA' profiles/$TAG/sass_$k.txt)"
done | tee profiles/$TAG/sass_summary.txt
