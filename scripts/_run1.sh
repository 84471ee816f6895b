timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_final.log 2>&1
tail -3 gpurun_out/r2_pytest_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke_final.log 2>&1
tail -2 gpurun_out/r2_smoke_final.log
timeout 1200 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
echo "bench rc=$?"
tail -c 400 gpurun_out/r2_bench_final.err
