mkdir -p gpurun_out
MM_NVCC_EXTRA="-DMM_WIDE_STAGES=4 -DMM_WIDE_SLABS=2" python -c "
import mm_s2ut_b200
from mm_s2ut_b200 import _lib
_lib.build(force=True)" > gpurun_out/rebuild.log 2>&1; tail -1 gpurun_out/rebuild.log
timeout 300 python profiles/tools/wide_sweep.py 2>&1 | tail -6
