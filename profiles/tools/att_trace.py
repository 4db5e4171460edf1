"""Per-phase clock64 timeline of the persistent attention kernel (build with MM_NVCC_EXTRA=-DMM_ATT_TRACE)."""
import ctypes
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import mm_s2ut_b200  # noqa: E402,F401
from mm_s2ut_b200 import _lib, kernels as K  # noqa: E402

dev = torch.device("cuda:0")
B, T, H, d = 64, 250, 8, 512
qkv = (torch.randn(B * T, 3 * d, device=dev) * 0.5).to(torch.bfloat16)
lens = torch.full((B,), T, dtype=torch.int32, device=dev)
out = torch.empty(B * T, d, dtype=torch.bfloat16, device=dev)
for _ in range(3):
    K.self_attention(qkv, lens, B, T, H, out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    K.self_attention(qkv, lens, B, T, H, out)
e1.record()
torch.cuda.synchronize()
print("kernel us", e0.elapsed_time(e1) * 100)
lib = _lib.load()
buf = np.zeros(148 * 2 * 8 * 6, dtype=np.int64)
lib.mm_debug_att_trace.restype = ctypes.c_int
lib.mm_debug_att_trace.argtypes = [ctypes.c_void_p]
print("rc", lib.mm_debug_att_trace(buf.ctypes.data))
t = buf.reshape(148, 2, 8, 6)
for cta in (0, 77, 147):
    base = t[cta][t[cta] > 0].min()
    for g in range(2):
        for it in range(4):
            r = t[cta, g, it]
            if r[0] == 0:
                continue
            print(f"cta {cta} g{g} item {it}: start {r[0]-base:7d} | wait S {r[1]-r[0]:6d} max {r[2]-r[1]:6d} exp {r[3]-r[2]:6d} "
                  f"wait O {r[4]-r[3]:6d} store {r[5]-r[4]:6d} | total {r[5]-r[0]:6d}")
d_ = t[:, :, 1:3, :]
d_ = d_[d_[..., 0] > 0]
names = ["wait S", "max", "exp", "wait O", "store"]
for k in range(5):
    print(f"mean {names[k]:7s} {np.mean(d_[:, k + 1] - d_[:, k]):9.0f} cycles")
