timeout 600 python -m pytest tests -m gpu -x -q -k "assembl or generations or box_mesh or distorted or quad8 or mid_size or reproduc" > gpurun_out/r2_pytest21.log 2>&1
tail -5 gpurun_out/r2_pytest21.log
timeout 300 python scripts/asm_probe.py 96 2 hexa8 "" "assembly_warp=0" > gpurun_out/r2_asm5.log 2>&1
timeout 300 python scripts/asm_probe.py 255 3 hexa8 "" "assembly_warp=0" >> gpurun_out/r2_asm5.log 2>&1
timeout 300 python scripts/asm_probe.py 94 3 hexa20 "" "assembly_warp=0" >> gpurun_out/r2_asm5.log 2>&1
cat gpurun_out/r2_asm5.log
