mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -4 gpurun_out/gpu_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench7.json 2> gpurun_out/bench7.err
echo "bench exit $?"; tail -5 gpurun_out/bench7.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench7.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['h2d_bytes_per_step'],'roof',d['roofline']['achieved'],d['roofline']['frac'])
for k,v in list(d['kernels'].items())[:6]: print(f"{k:22s} n={v['launches_per_step']:3d} ms={v['ms_per_step']:.4f} share={v['share']:.3f} ach={v['achieved']:.1f} {v['unit']}")
PY
