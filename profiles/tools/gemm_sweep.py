"""GEMM micro-benchmark: per-tile period vs K for each epilogue mode (separates epilogue cost from the
per-k-block mainloop rate).  M = 74 pairs x 256 rows, N = 2048 -> exactly 8 tiles per CTA pair."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import mm_s2ut_b200  # noqa: E402,F401
from mm_s2ut_b200 import kernels as K  # noqa: E402

dev = torch.device("cuda:0")
M, N = 74 * 256, 2048
dt = torch.bfloat16


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


bias = torch.randn(N, device=dev)
out16 = torch.empty(M, N, dtype=dt, device=dev)
out32 = torch.empty(M, N, dtype=torch.float32, device=dev)
x32 = torch.randn(M, N, device=dev)
print(f"M={M} N={N}: 8 tiles per CTA pair; period = time/8; cycles at 1.9 GHz")
for name, mode in (("relu_op", K.EPI_RELU_OP), ("op", K.EPI_OP), ("f32", K.EPI_F32), ("resid_f32", K.EPI_RESID_F32)):
    for Kd in (64, 256, 512, 1024, 2048, 4096):
        a = torch.randn(M, Kd, device=dev).to(dt)
        w = (torch.randn(N, Kd, device=dev) * Kd ** -0.5).to(dt)
        if mode == K.EPI_RESID_F32:
            fn = lambda: K.gemm(a0=a, a0_ld=Kd, rows=M, w=w, n=N, k=Kd, mode=mode, bias=bias, aux0=x32, aux_ld=N,
                                out0=out32, out0_ld=N)
        elif mode == K.EPI_F32:
            fn = lambda: K.gemm(a0=a, a0_ld=Kd, rows=M, w=w, n=N, k=Kd, mode=mode, bias=bias, out0=out32, out0_ld=N)
        else:
            fn = lambda: K.gemm(a0=a, a0_ld=Kd, rows=M, w=w, n=N, k=Kd, mode=mode, bias=bias, out0=out16, out0_ld=N)
        us = timeit(fn)
        tf = 2.0 * M * N * Kd / us / 1e6
        print(f"{name:10s} K={Kd:5d}  {us:8.1f} us  {tf:7.1f} TFLOP/s  period/tile {us / 8:6.2f} us = "
              f"{us / 8 * 1900:7.0f} cyc  per k-block {us / 8 * 1900 / (Kd / 64):6.0f} cyc")
