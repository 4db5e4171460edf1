"""Run the fbank kernel alone at the bench shape (B = 64 x 10 s) -- for `ncu -k regex:fbank_kernel`."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import mm_s2ut_b200  # noqa: E402,F401
from mm_s2ut_b200 import kernels as K  # noqa: E402

dev = torch.device("cuda:0")
B, n = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 160000
m = 1 + (n - 400) // 160
wav = (torch.randn(B, n, device=dev) * 3000).round()
if len(sys.argv) > 2 and sys.argv[2] == "i16":
    wav = wav.to(torch.int16)
lens = torch.full((B,), n, dtype=torch.int64, device=dev)
raw = torch.empty(B, m, 80, device=dev)
tables = K.fbank_tables(dev)
for _ in range(3):
    K.fbank(wav, lens, raw, tables)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    K.fbank(wav, lens, raw, tables)
e1.record()
torch.cuda.synchronize()
print(f"fbank B={B} {wav.dtype}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
