python scripts/asm_probe.py 255 3 hexa8 "" "assembly_records=0" > gpurun_out/r2_asm2.log 2>&1
python scripts/asm_probe.py 94 3 hexa20 "" >> gpurun_out/r2_asm2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_assemble|k_elem_records|k_blk_desc" --csv --log-file gpurun_out/r2_asm_launches.csv python scripts/asm_probe.py 128 2 hexa8 "" "assembly_records=0" > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_assemble_rec" -c 1 -o gpurun_out/r2_k_assemble_rec_128cube python scripts/asm_probe.py 128 1 hexa8 > /dev/null 2>&1
cat gpurun_out/r2_asm2.log
grep -v "^==" gpurun_out/r2_asm_launches.csv | cut -d, -f5,12,13,15 | head -20
