set -x
ORDER20=interleaved PARITY=1 python scripts/precond_probe.py 94 20 hexa20 "" "fsai_vertex_first=0" > gpurun_out/r2_precond10.log 2>&1
ORDER20=grouped PARITY=1 python scripts/precond_probe.py 94 20 hexa20 "" >> gpurun_out/r2_precond10.log 2>&1
python -m pytest tests -m gpu -x -q -k "hexa20 or newmark or bathe or fsai or precond" > gpurun_out/r2_pytest16.log 2>&1
tail -3 gpurun_out/r2_pytest16.log
cat gpurun_out/r2_precond10.log
