mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rowwise.py -q -m gpu --timeout 120 -k "attention" > gpurun_out/attn.log 2>&1
echo "attn exit $?"; grep -E "passed|failed|Error|timeout|assert [0-9]|^E  |mbarrier" gpurun_out/attn.log | head -30
timeout 600 python -m pytest tests/test_gpu_encoder.py -q -m gpu --timeout 120 > gpurun_out/enc.log 2>&1
echo "enc exit $?"; tail -3 gpurun_out/enc.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench5.json 2> gpurun_out/bench5.err
echo "bench exit $?"; tail -5 gpurun_out/bench5.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench5.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'roof',d['roofline']['achieved'],d['roofline']['frac'])
for k,v in d['kernels'].items(): print(f"{k:22s} n={v['launches_per_step']:3d} ms={v['ms_per_step']:.4f} share={v['share']:.3f} ach={v['achieved']:.1f} {v['unit']}")
PY
