mkdir -p gpurun_out
for f in test_gpu_rowwise test_gpu_encoder; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu --timeout 300 > gpurun_out/$f.log 2>&1
  echo "$f exit $?"
  grep -E "passed|failed|Error|assert [0-9]|^E  " gpurun_out/$f.log | head -40
done
