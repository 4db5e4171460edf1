"""Times mm_colsum / mm_reduce_partials_many in isolation (CUDA events, L2 flushed) on zero and random data."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import mm_s2ut_b200  # noqa
from mm_s2ut_b200 import kernels as K

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=10):
    ts = []
    for i in range(n + 2):
        flush.fill_(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return min(ts[2:]), sum(ts[2:]) / n


M = 16000
for cols in (512, 1536, 2048):
    for name, x in (("zeros", torch.zeros(M, cols, dtype=torch.bfloat16, device=dev)),
                    ("randn", torch.randn(M, cols, device=dev).bfloat16())):
        nb = K.colsum_blocks(M)
        part = torch.empty(nb * cols, device=dev)
        out = torch.empty(cols, device=dev)
        b, a = timeit(lambda: K.colsum(x, cols, M, cols, part))
        b2, a2 = timeit(lambda: K.reduce_partials_many([(part, nb, cols, cols, out, False)]))
        b3, a3 = timeit(lambda: x.float().sum(0))
        print(f"cols {cols:5d} {name}: colsum best {b:6.1f} avg {a:6.1f} us ({2.0 * M * cols / b / 1e3:6.0f} GB/s)   "
              f"reduce({nb} partials) best {b2:5.1f} us   torch sum {b3:6.1f} us")
# the split-K partial reduction of a weight gradient: 9 partials of 2048 x 512
n = 2048 * 512
for S in (4, 9, 18):
    part = torch.randn(S * n, device=dev)
    out = torch.empty(n, device=dev)
    b, a = timeit(lambda: K.reduce_partials_many([(part, S, n, n, out, False)]))
    print(f"reduce S={S} n={n}: best {b:6.1f} avg {a:6.1f} us ({4.0 * n * (S + 1) / b / 1e3:6.0f} GB/s)")
