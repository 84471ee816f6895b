timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest20.log 2>&1
tail -3 gpurun_out/r2_pytest20.log
timeout 1200 python bench.py > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
tail -c 600 gpurun_out/r2_bench_b.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --stage 20 --no-cpu --parity 0 --e2e-sweep 0 --scatter-e2e 0 --secondary 0 --random-field 0 > gpurun_out/r2_bench_ncu.log 2>&1
T="tests/test_gpu_parity.py::test_assembly_kernel_generations_agree tests/test_gpu_parity.py::test_box_mesh_random_field tests/test_gpu_parity.py::test_thin_slab_with_tiles_of_ghost_nodes"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest $T -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2_sanitize_asm.log
echo "exit: $?" >> gpurun_out/r2_sanitize_asm.log
cat gpurun_out/r2_sanitize_asm.log
