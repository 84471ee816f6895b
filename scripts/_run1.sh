timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_final3.log 2>&1
tail -3 gpurun_out/r2_pytest_final3.log
timeout 200 python scripts/asm_probe.py 255 3 hexa8 > gpurun_out/r2_asm14.log 2>&1
timeout 200 python scripts/asm_probe.py 94 2 hexa20 >> gpurun_out/r2_asm14.log 2>&1
cat gpurun_out/r2_asm14.log
