timeout 600 python -m pytest tests -m gpu -x -q -k "assembl or generations or box_mesh or distorted or quad8 or mid_size or reproduc or hexa20 or thin_slab" > gpurun_out/r2_pytest26.log 2>&1
tail -3 gpurun_out/r2_pytest26.log
timeout 300 python scripts/asm_probe.py 255 3 hexa8 "" > gpurun_out/r2_asm11.log 2>&1
timeout 300 python scripts/asm_probe.py 94 3 hexa20 "" >> gpurun_out/r2_asm11.log 2>&1
cat gpurun_out/r2_asm11.log
