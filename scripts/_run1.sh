timeout 600 python -m pytest tests -m gpu -x -q -k "assembl or generations or box_mesh or distorted or quad8 or mid_size or reproduc" > gpurun_out/r2_pytest19.log 2>&1
tail -3 gpurun_out/r2_pytest19.log
timeout 300 python scripts/asm_probe.py 255 3 hexa8 "" > gpurun_out/r2_asm4.log 2>&1
timeout 300 python scripts/asm_probe.py 94 3 hexa20 "" >> gpurun_out/r2_asm4.log 2>&1
cat gpurun_out/r2_asm4.log
ncu --set full --clock-control none --import-source on -k regex:"k_assemble_tma" -c 1 -o gpurun_out/r2_k_assemble_tma_128cube_b python scripts/asm_probe.py 128 1 hexa8 > /dev/null 2>&1
