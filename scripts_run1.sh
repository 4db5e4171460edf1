mkdir -p gpurun_out
rm -f gpurun_out/parity.txt
MM_PARITY_REPORT=$PWD/gpurun_out/parity.txt timeout 900 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -3 gpurun_out/gpu_tests.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench exit $?"; tail -2 gpurun_out/bench_default.err
python -c "
import json
d=json.load(open('gpurun_out/bench_default.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'roof',d['roofline']['frac'],d['clocks'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_v18.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/launches_v18.csv
