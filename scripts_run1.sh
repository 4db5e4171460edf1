mkdir -p gpurun_out
timeout 300 python profiles/tools/att_trace.py 2>&1 | tail -32
timeout 600 python -m pytest tests/test_gpu_rowwise.py tests/test_gpu_encoder.py -q -m gpu --timeout 300 -x > gpurun_out/att_tests.log 2>&1
echo "tests exit $?"; tail -4 gpurun_out/att_tests.log
