"""mm_gemm_ln_bwd against the un-fused pair (mm_gemm EPI_F32 + mm_layernorm_bwd) at the training shapes: device time per
call, 30 back-to-back calls between two events (launch latency hidden behind the queue)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import mm_s2ut_b200  # noqa: E402,F401
from mm_s2ut_b200 import kernels as K  # noqa: E402

dev = torch.device("cuda:0")
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 16000
x = torch.randn(rows, 512, device=dev)
gamma = torch.ones(512, device=dev)
g = torch.randn(rows, 512, device=dev) * 0.1
g_op = torch.empty(rows, 512, dtype=torch.bfloat16, device=dev)
dh = torch.empty(rows, 512, device=dev)
part = torch.empty(max(K.gemm_ln_bwd_partial_rows(rows), K.layernorm_bwd_blocks() * 2) * 1024, device=dev)


def timed(fn, n=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for k in (64, 512, 1536, 2048):
    dy = (torch.randn(rows, k, device=dev) * 0.05).bfloat16()
    w = (torch.randn(k, 512, device=dev) * 0.05).bfloat16()
    fused = timed(lambda: K.gemm_ln_bwd(dy, w, x, gamma, g, g_op, part))

    def pair():
        K.gemm(a0=dy, a0_ld=k, rows=rows, w=w, w_ld=512, w_mn=True, n=512, k=k, mode=K.EPI_F32, out0=dh, out0_ld=512)
        K.layernorm_bwd(x, gamma, dh, part, dx=g, resid=g, dx_op=g_op)

    unfused = timed(pair)
    gemm_only = timed(lambda: K.gemm(a0=dy, a0_ld=k, rows=rows, w=w, w_ld=512, w_mn=True, n=512, k=k, mode=K.EPI_F32,
                                     out0=dh, out0_ld=512))
    print(f"rows {rows} k {k:5d}: fused {fused:6.1f} us   gemm + ln_bwd {unfused:6.1f} us   (gemm alone {gemm_only:5.1f})", flush=True)
