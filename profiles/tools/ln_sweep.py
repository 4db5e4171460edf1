"""Fused GEMM+residual+LayerNorm micro-benchmark: time vs K at one tile per CTA pair (M = 74 x 256)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import mm_s2ut_b200  # noqa: E402,F401
from mm_s2ut_b200 import kernels as K  # noqa: E402

dev = torch.device("cuda:0")
N, dt = 512, torch.bfloat16


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
    return e0.elapsed_time(e1) / iters * 1e3


for M in (74 * 256, 16000):
    bias, g, b = torch.randn(N, device=dev), torch.randn(N, device=dev), torch.randn(N, device=dev)
    x = torch.randn(M, N, device=dev)
    h = torch.empty(M, N, dtype=dt, device=dev)
    for Kd in (64, 512, 2048, 4096):
        a = torch.randn(M, Kd, device=dev).to(dt)
        w = (torch.randn(N, Kd, device=dev) * Kd ** -0.5).to(dt)
        us = timeit(lambda: K.gemm_resid_ln(a, w, bias, x, g, b, h))
        us2 = timeit(lambda: (K.gemm(a0=a, a0_ld=Kd, rows=M, w=w, n=N, k=Kd, mode=K.EPI_RESID_F32, bias=bias, aux0=x,
                                     aux_ld=N, out0=x, out0_ld=N), K.layernorm(x, g, b, out_op=h)))
        print(f"M={M} K={Kd:5d} fused {us:7.1f} us ({us * 1900:8.0f} cyc)   unfused gemm+ln {us2:7.1f} us")
