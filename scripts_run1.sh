mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -4 gpurun_out/gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench exit $?"; tail -3 gpurun_out/bench_default.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_default.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks','cpu_baseline')}); print(d['e2e']); print(d['roofline']['achieved'], d['roofline']['frac'])
PY
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "ref exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_ref.json')); print(d['value'], d['cpu_baseline']['cores'], d['ms_per_step'])"
