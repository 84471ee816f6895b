timeout 300 python scripts/asm_probe.py 96 2 hexa8 "" "assembly_records=0" > gpurun_out/r2_asm9.log 2>&1
timeout 300 python scripts/asm_probe.py 255 3 hexa8 "" >> gpurun_out/r2_asm9.log 2>&1
cat gpurun_out/r2_asm9.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_elem_records|k_assemble_tma" --csv --log-file gpurun_out/r2_asm_launches3.csv python scripts/asm_probe.py 255 1 hexa8 > /dev/null 2>&1
grep -v "^==" gpurun_out/r2_asm_launches3.csv | tail -2 | cut -c1-90,380-
