#!/bin/bash
# compute-sanitizer passes over the kernels (the equivalent of the reference's race/consistency checks,
# SURVEY.md section 5).  Run under gpurun: gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
set -u
OUT=gpurun_out
mkdir -p $OUT
TESTS="tests/test_gpu_parity.py::test_fixture_sigs_and_manysearch tests/test_gpu_parity.py::test_coverage_scaled50_with_n_run tests/test_gpu_parity.py::test_exact_tile_multiple_and_boundaries tests/test_gpu_parity.py::test_intersect_segmentation_variants tests/test_gpu_parity.py::test_intersect_rectangular_and_sharded tests/test_gpu_parity.py::test_scaled_one_keeps_everything tests/test_gpu_parity.py::test_intersect_index_form_equals_probe_and_oracle tests/test_gpu_parity.py::test_survivor_workspace_and_direct_insert_agree tests/test_gpu_parity.py::test_ingest_pipeline_paths_agree tests/test_gpu_parity.py::test_step_pipeline_eager_and_graph"
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --target-processes all --log-file $OUT/sanitizer_$tool.log \
      python -m pytest $TESTS -m gpu -x -q > $OUT/sanitizer_${tool}_pytest.log 2>&1
  echo "$tool: pytest rc=$? ; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/sanitizer_$tool.log | tail -1)"
  tail -1 $OUT/sanitizer_${tool}_pytest.log
done
