mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -3 gpurun_out/gpu_tests.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench17.json 2> gpurun_out/bench17.err
echo "bench exit $?"; tail -3 gpurun_out/bench17.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench17.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'roof',d['roofline']['achieved'],d['roofline']['frac'])
for k,v in d['kernels'].items(): print(k, round(v['ms_per_step'],4), v['launches_per_step'])
PY
