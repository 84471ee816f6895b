#!/bin/bash
# compute-sanitizer pass over a small slice of the GPU suite (memcheck + racecheck); run on the GPU box:
#   bash scripts/sanitize.sh > gpurun_out/sanitize.log 2>&1
set -u
T="tests/test_gpu_parity.py::test_assembly_matches_oracle_and_reference[column_abs] tests/test_gpu_parity.py::test_assembly_matches_oracle_and_reference[column_2D_tri6] tests/test_gpu_parity.py::test_assembly_matches_oracle_and_reference[column_high_order] tests/test_gpu_parity.py::test_quad8_elements tests/test_gpu_parity.py::test_rows_without_entries_do_not_disturb_pcg tests/test_gpu_parity.py::test_newmark_quad4_heaviside_vs_reference_golden tests/test_gpu_parity.py::test_central_difference_vs_oracle tests/test_gpu_parity.py::test_box_mesh_random_field"
for tool in memcheck racecheck; do
  echo "=== compute-sanitizer --tool $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 5 python -m pytest $T -m gpu -q -x 2>&1 | tail -15
  echo "exit: $?"
done
