#!/bin/bash
# 2-GPU check: multi-GPU CUDA tests + the bench under torchrun
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_multi_gpu_cuda.py -x -q > $OUT/r2m_tests.log 2>&1; echo "rc=$?" >> $OUT/r2m_tests.log
tail -5 $OUT/r2m_tests.log
bash tools/gpu_scale.sh r2m config3
