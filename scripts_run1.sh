mkdir -p gpurun_out
MM_GEMM_DIRECT=1 timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_encoder.py -q -m gpu --timeout 120 -x > gpurun_out/direct_tests.log 2>&1
echo "tests exit $?"; tail -3 gpurun_out/direct_tests.log
for d in 1 0 1 0; do
echo "== direct $d"
MM_GEMM_DIRECT=$d timeout 300 python profiles/tools/gemm_sweep.py 2>&1 | grep -E "^relu_op" | grep -E "K=  512|K= 2048"
MM_GEMM_DIRECT=$d timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms',d['ms_per_step'], 'op', d['kernels']['gemm[op]']['ms_per_step'], 'relu', d['kernels']['gemm[relu_op]']['ms_per_step'])"
done
