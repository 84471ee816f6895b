#!/bin/bash
# compute-sanitizer over the kernels added in round 2 (FSAI set-up / products, projection, two-group ring of k_spmv_node,
# ghost-only tiles); run on the GPU box:   bash scripts/sanitize_r2.sh > gpurun_out/sanitize_r2.log 2>&1
set -u
T="tests/test_gpu_parity.py::test_newmark_quad4_heaviside_vs_reference_golden[graph] tests/test_gpu_parity.py::test_newmark_quad4_heaviside_vs_reference_golden[eager_jacobi_projection] tests/test_gpu_parity.py::test_thin_slab_with_tiles_of_ghost_nodes tests/test_gpu_parity.py::test_bathe_and_static_vs_oracle[True] tests/test_gpu_parity.py::test_rows_without_entries_do_not_disturb_pcg[True] tests/test_gpu_parity.py::test_newmark_absorbing_and_hexa20_vs_oracle[True]"
for tool in ${TOOLS:-memcheck racecheck}; do
  echo "=== compute-sanitizer --tool $tool"
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 5 python -m pytest $T -m gpu -q -x 2>&1 | tail -12
  echo "exit: $?"
done
