mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rowwise.py tests/test_gpu_int16.py tests/test_gpu_encoder.py -q -m gpu --timeout 300 > gpurun_out/t.log 2>&1
echo "tests exit $?"; grep -E "passed|failed|Error|timeout|assert [0-9]|^E  " gpurun_out/t.log | head
python profiles/tools/frontend_sweep.py 2>&1 | grep -E "float32" | tee gpurun_out/frontend_sweep_v2.txt
