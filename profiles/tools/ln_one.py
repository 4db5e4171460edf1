"""One shape of the fused GEMM+residual+LN kernel (for ncu): python ln_one.py M K [iters]."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import mm_s2ut_b200  # noqa: E402,F401
from mm_s2ut_b200 import kernels as K  # noqa: E402

M, Kd = int(sys.argv[1]), int(sys.argv[2])
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev, N, dt = torch.device("cuda:0"), 512, torch.bfloat16
bias, g, b = torch.randn(N, device=dev), torch.randn(N, device=dev), torch.randn(N, device=dev)
x = torch.randn(M, N, device=dev)
h = torch.empty(M, N, dtype=dt, device=dev)
a = torch.randn(M, Kd, device=dev).to(dt)
w = (torch.randn(N, Kd, device=dev) * Kd ** -0.5).to(dt)
for _ in range(iters):
    K.gemm_resid_ln(a, w, bias, x, g, b, h)
    K.gemm(a0=a, a0_ld=Kd, rows=M, w=w, n=N, k=Kd, mode=K.EPI_RESID_F32, bias=bias, aux0=x, aux_ld=N, out0=x, out0_ld=N)
torch.cuda.synchronize()
print("ok")
