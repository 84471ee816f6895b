#!/bin/bash
# compute-sanitizer pass over a small slice of the GPU suite (memcheck + racecheck); run on the GPU box:
#   bash scripts/sanitize.sh > gpurun_out/sanitize.log 2>&1
set -u
T="tests/test_gpu_parity.py::test_assembly_matches_oracle_and_reference[column_abs] tests/test_gpu_parity.py::test_assembly_matches_oracle_and_reference[column_2D_tri6] tests/test_gpu_parity.py::test_assembly_matches_oracle_and_reference[column_high_order] tests/test_gpu_parity.py::test_quad8_elements tests/test_gpu_parity.py::test_rows_without_entries_do_not_disturb_pcg tests/test_gpu_parity.py::test_newmark_quad4_heaviside_vs_reference_golden tests/test_gpu_parity.py::test_central_difference_vs_oracle tests/test_gpu_parity.py::test_box_mesh_random_field"
# kernels added later in round 1: column dictionary (node_dict.cu + k_spmv_node), random field (k_srf), absorbing faces (absorb.cu)
T="$T tests/test_gpu_parity.py::test_column_dictionary_is_bitwise_neutral[cube_abs] tests/test_gpu_parity.py::test_column_dictionary_is_bitwise_neutral[rose_2D_side] tests/test_gpu_parity.py::test_random_field_kernel_vs_oracle[Gaussian] tests/test_gpu_parity.py::test_scatter_with_random_field tests/test_gpu_parity.py::test_assembly_matches_oracle_and_reference[cube_abs] tests/test_gpu_parity.py::test_newmark_absorbing_and_hexa20_vs_oracle"
for tool in ${TOOLS:-memcheck racecheck}; do
  echo "=== compute-sanitizer --tool $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 5 python -m pytest $T -m gpu -q -x 2>&1 | tail -15
  echo "exit: $?"
done
