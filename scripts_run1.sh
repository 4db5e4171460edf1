mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_configs.py -q -m gpu --timeout 300 > gpurun_out/cfg_tests.log 2>&1
echo "cfg tests exit $?"; grep -E "passed|failed|Error|timeout|assert [0-9]|^E  |mbarrier" gpurun_out/cfg_tests.log | head -30
