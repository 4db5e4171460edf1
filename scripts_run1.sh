mkdir -p gpurun_out
for f in test_gpu_rowwise test_gpu_encoder; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu --timeout 300 > gpurun_out/$f.log 2>&1
  echo "$f exit $?"; tail -3 gpurun_out/$f.log
done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err
echo "bench exit $?"; tail -5 gpurun_out/bench1.err; cat gpurun_out/bench1.json
