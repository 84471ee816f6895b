timeout 140 python bench.py --steps 3 --warmup 3 --no-cpu --e2e-sweep 0 --scatter-e2e 0 --secondary 0 --random-field 0 > gpurun_out/r2_bench_final_short.json 2> gpurun_out/r2_bench_final_short.err
echo "rc=$?"; tail -c 300 gpurun_out/r2_bench_final_short.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2_bench_final_short.json') if x.startswith('{')]
d=json.loads(l[-1]); print(d['value'], d['e2e']['value'], d['assembly']['seconds'], d['assembly']['pattern_seconds'], d['parity_check']['passed'], d['roofline']['frac'])
PY
