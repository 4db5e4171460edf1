"""Actual SM clock under load: clock64 vs globaltimer inside the persistent GEMM (build with -DMM_GEMM_TRACE)."""
import ctypes
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import mm_s2ut_b200  # noqa: E402,F401
from mm_s2ut_b200 import _lib, kernels as K  # noqa: E402

dev = torch.device("cuda:0")
M, N, dt = 74 * 256, 2048, torch.bfloat16
lib = _lib.load()
lib.mm_debug_gemm_trace.restype = ctypes.c_int
lib.mm_debug_gemm_trace.argtypes = [ctypes.c_void_p]
out = torch.empty(M, N, dtype=dt, device=dev)
bias = torch.randn(N, device=dev)
for Kd in (512, 2048, 4096):
    for kind in ("randn", "zeros"):
        a = (torch.randn(M, Kd, device=dev) if kind == "randn" else torch.zeros(M, Kd, device=dev)).to(dt)
        w = ((torch.randn(N, Kd, device=dev) * Kd ** -0.5) if kind == "randn" else torch.zeros(N, Kd, device=dev)).to(dt)
        for _ in range(20):
            K.gemm(a0=a, a0_ld=Kd, rows=M, w=w, n=N, k=Kd, mode=K.EPI_RELU_OP, bias=bias, out0=out, out0_ld=N)
        torch.cuda.synchronize()
        buf = np.zeros(148 * 4, dtype=np.int64)
        assert lib.mm_debug_gemm_trace(buf.ctypes.data) == 0
        t = buf.reshape(148, 4)
        cyc = (t[:, 1] - t[:, 0]).astype(np.float64)
        ns = (t[:, 3] - t[:, 2]).astype(np.float64)
        kb = 8 * Kd / 64
        print(f"K={Kd:5d} {kind:6s}: in-kernel {cyc.mean():9.0f} cycles = {ns.mean() / 1e3:7.1f} us -> SM clock {cyc.mean() / ns.mean():.3f} GHz;"
              f" {cyc.mean() / kb:6.0f} cycles per k-block incl. epilogue tail")
