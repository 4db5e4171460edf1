mkdir -p gpurun_out
for r in 1 0 1 0; do
MM_LN_RECUT=$r timeout 600 python bench.py --steps 300 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('recut $r value',d['value'],'ms',d['ms_per_step'], d['clocks'])"
done
