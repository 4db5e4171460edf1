"""Self-attention core across sequence lengths: T <= 256 takes the persistent TMEM kernel, longer sequences the
general chunked kernel.  ~16 k tokens per call, 8 heads of 64."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import mm_s2ut_b200  # noqa: E402,F401
from mm_s2ut_b200 import kernels as K  # noqa: E402

dev = torch.device("cuda:0")
H, d = 8, 512


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


for T in (125, 250, 256, 300, 500, 750):
    B = max(1, 16000 // T)
    qkv = (torch.randn(B * T, 3 * d, device=dev) * 0.5).to(torch.bfloat16)
    lens = torch.full((B,), T, dtype=torch.int32, device=dev)
    out = torch.empty(B * T, d, dtype=torch.bfloat16, device=dev)
    us = timeit(lambda: K.self_attention(qkv, lens, B, T, H, out))
    fl = 4.0 * B * H * T * T * 64
    print(f"T={T:4d} B={B:3d}: {us:7.1f} us  {fl / us / 1e6:6.1f} TFLOP/s (useful)  {B * T / us:7.1f} tokens/us")
