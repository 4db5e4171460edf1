"""Phase timeline (clock64) of the fused GEMM+residual+LayerNorm kernel inside a realistic layer sequence
(QKV GEMM -> attention -> out_proj+LN -> FFN1 -> FFN2+LN at the bench shape), so L2 holds what it would hold in
the encoder.  Build with MM_NVCC_EXTRA=-DMM_LN_TRACE."""
import ctypes
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import mm_s2ut_b200  # noqa: E402,F401
from mm_s2ut_b200 import _lib, kernels as K  # noqa: E402

dev = torch.device("cuda:0")
B, T, H, d, F = 64, 250, 8, 512, 2048
M, dt = B * T, torch.bfloat16
lib = _lib.load()
lib.mm_debug_ln_trace.restype = ctypes.c_int
lib.mm_debug_ln_trace.argtypes = [ctypes.c_void_p]
r = lambda *s: torch.randn(*s, device=dev)
wqkv, bqkv = (r(3 * d, d) * d ** -0.5).to(dt), r(3 * d)
wo, bo = (r(d, d) * d ** -0.5).to(dt), r(d)
w1, b1 = (r(F, d) * d ** -0.5).to(dt), r(F)
w2, b2 = (r(d, F) * F ** -0.5).to(dt), r(d)
g, be = torch.ones(d, device=dev), torch.zeros(d, device=dev)
x = r(M, d)
h = r(M, d).to(dt)
qkv = torch.empty(M, 3 * d, dtype=dt, device=dev)
att = torch.empty(M, d, dtype=dt, device=dev)
f = torch.empty(M, F, dtype=dt, device=dev)
lens = torch.full((B,), T, dtype=torch.int32, device=dev)


def layer(stop_after_out_proj=False):
    K.gemm(a0=h, a0_ld=d, rows=M, w=wqkv, n=3 * d, k=d, mode=K.EPI_OP, bias=bqkv, scale=0.125, scale_cols=d, out0=qkv,
           out0_ld=3 * d)
    K.self_attention(qkv, lens, B, T, H, att)
    K.gemm_resid_ln(att, wo, bo, x, g, be, h)
    if stop_after_out_proj:
        return
    K.gemm(a0=h, a0_ld=d, rows=M, w=w1, n=F, k=d, mode=K.EPI_RELU_OP, bias=b1, out0=f, out0_ld=F)
    K.gemm_resid_ln(f, w2, b2, x, g, be, h)


names = ["mainloop (start -> acc ready)", "sweep 1 (residual in/out)", "stat exchange",
         "sweep 2 (normalise, 16-bit out)", "drain + cluster sync", "exit wait"]
for which in ("out_proj K=512", "fc2 K=2048"):
    for _ in range(3):
        layer()
    layer(stop_after_out_proj=which.startswith("out"))
    torch.cuda.synchronize()
    buf = np.zeros(148 * 2 * 16, dtype=np.int64)
    assert lib.mm_debug_ln_trace(buf.ctypes.data) == 0
    t = buf.reshape(148, 2, 16)[:126]
    print(f"{which}: mean cycles over {t.shape[0]} CTAs x 2 halves")
    for k in range(6):
        dlt = t[:, :, k + 1] - t[:, :, k]
        print(f"  {names[k]:34s} mean {dlt.mean():8.0f}  min {dlt.min():8.0f}  max {dlt.max():8.0f}")
    print(f"  total traced {np.mean(t[:, :, 6] - t[:, :, 0]):8.0f} cycles")
    sub = ["step-1 start -> wait_read+bar", "tmem wait + issue", "normalise cols 0-31", "tmem wait + normalise cols 32-63",
           "fence + bar"]
    for k in range(5):
        dlt = t[:, :, 8 + k] - t[:, :, 7 + k]
        print(f"    sweep-2 step 1: {sub[k]:34s} mean {dlt.mean():7.0f}  max {dlt.max():7.0f}")

# hot-L2 bound: the same kernel back to back (x, a, h all resident)
for Kd, a_, w_, b_ in ((512, att, wo, bo), (2048, f, w2, b2)):
    for _ in range(4):
        K.gemm_resid_ln(a_, w_, b_, x, g, be, h)
    torch.cuda.synchronize()
    buf = np.zeros(148 * 2 * 16, dtype=np.int64)
    assert lib.mm_debug_ln_trace(buf.ctypes.data) == 0
    t = buf.reshape(148, 2, 16)[:126]
    print(f"back-to-back K={Kd} (everything L2-resident):")
    for k in range(5):
        dlt = t[:, :, k + 1] - t[:, :, k]
        print(f"  {names[k]:34s} mean {dlt.mean():8.0f}")
