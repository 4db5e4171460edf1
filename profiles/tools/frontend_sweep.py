"""BASELINE configs[3]: fbank + CMVN + Conv1dSubsampler front-end throughput sweep, batch 256, 1-30 s utterances."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import mm_s2ut_b200  # noqa: E402,F401
from mm_s2ut_b200 import kernels as K  # noqa: E402
from mm_s2ut_b200.config import make_args  # noqa: E402
from mm_s2ut_b200.models.mm_s2s_transformer import MM_S2STransformerEncoder  # noqa: E402

HBM = 6454.9
dev = torch.device("cuda:0")
torch.manual_seed(0)
enc = MM_S2STransformerEncoder(make_args("base"), build_unused_projections=False).eval().to(dev)
eng = enc.engine()
B = 256


def timeit(fn, iters=10):
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


print(f"B={B}; fbank algorithmic bytes = 4*n + 320*m per utterance; HBM peak {HBM} GB/s (measured copy)")
for dur in (1, 2, 5, 10, 20, 30):
    n = dur * 16000
    m = 1 + (n - 400) // 160
    for dt in (torch.float32, torch.int16):
        wav = (torch.randn(B, n, device=dev) * 3000).round().to(dt)
        lens = torch.full((B,), n, dtype=torch.int64, device=dev)
        raw = torch.empty(B, m, 80, device=dev)
        ms = torch.empty(B, 2, 80, device=dev)
        us_fb = timeit(lambda: K.fbank(wav, lens, raw, eng.fbank_tables))
        us_st = timeit(lambda: K.cmvn_stats(raw, lens, True, ms))
        us_all = timeit(lambda: eng.subsample(*eng.frontend(wav, lens)[:3]))
        bytes_fb = B * (wav.element_size() * n + 320 * m)
        print(f"{dur:2d} s {str(dt)[6:]:8s} fbank {us_fb:8.1f} us = {bytes_fb / us_fb / 1e3:7.1f} GB/s ({bytes_fb / us_fb / 1e3 / HBM:5.1%} of HBM)"
              f"  cmvn_stats {us_st:7.1f} us  front-end (fbank+cmvn+2 conv) {us_all:8.1f} us = {B * dur / us_all * 1e6:10.0f} audio-s/s")
