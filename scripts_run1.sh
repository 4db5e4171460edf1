mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -3 gpurun_out/gpu_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench8.json 2> gpurun_out/bench8.err
echo "bench exit $?"; tail -5 gpurun_out/bench8.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench8.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'roof',d['roofline']['achieved'],d['roofline']['frac'])
for k,v in list(d['kernels'].items()): print(f"{k:22s} n={v['launches_per_step']:3d} ms={v['ms_per_step']:.4f} share={v['share']:.3f} ach={v['achieved']:.1f} {v['unit']}")
PY
python profiles/tools/gemm_sweep.py 2>&1 | grep -E "K=   64|K=  512|K= 2048" | tee gpurun_out/gemm_sweep_v3.txt
python profiles/tools/ln_sweep.py 2>&1 | tee gpurun_out/ln_sweep_v4.txt
