"""NCCL all-reduce of the 162 MB fp32 gradient buffer: bucket size and dtype sweep (device-timed, max over ranks).
torchrun --nproc-per-node N profiles/tools/allreduce_probe.py"""
import os

import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
n = 42701824
flat = torch.randn(n, device="cuda")
half = flat.bfloat16()


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def bucketed(buf, be):
    def f():
        works = [dist.all_reduce(buf[o:o + be], async_op=True) for o in range(0, buf.numel(), be)]
        for w in works:
            w.wait()
    return f


for name, buf in (("fp32", flat), ("bf16", half)):
    for be in (2 << 20, 8 << 20, 16 << 20, n):
        ms = timed(bucketed(buf, be))
        nbytes = buf.numel() * buf.element_size()
        if rank == 0:
            print(f"{name} bucket {be * buf.element_size() >> 20:4d} MB: {ms:.3f} ms  busbw "
                  f"{2 * (world - 1) / world * nbytes / ms / 1e6:.0f} GB/s", flush=True)
ms = timed(lambda: (half.copy_(flat), dist.all_reduce(half), flat.copy_(half)))
if rank == 0:
    print(f"fp32 -> bf16 convert + all-reduce + back: {ms:.3f} ms")
# ---- the peer-memory kernel (csrc/p2p.cu) against NCCL: same sums, device time incl. its two barriers
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import mm_s2ut_b200  # noqa: E402,F401
from mm_s2ut_b200.peer import PeerGroup  # noqa: E402

for wire in ("fp32", "bf16"):
    torch.manual_seed(rank)
    a = torch.randn(n, device="cuda")
    b = a.clone()
    dist.all_reduce(b)
    grp = PeerGroup(a, exchange_dtype=wire)
    grp.all_reduce()
    torch.cuda.synchronize()
    err = (a - b).abs().max().item()
    rel = ((a - b).norm() / b.norm()).item()
    same = torch.tensor([float(a.double().sum().item())], device="cuda", dtype=torch.float64)
    lo, hi = same.clone(), same.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    ms = timed(grp.all_reduce)
    if rank == 0:
        nbytes = n * 4
        what = "barrier + kernel + barrier" if wire == "fp32" else "pack + barrier + kernel + barrier + unpack"
        print(f"peer-memory all-reduce, {wire} on the wire ({what}): {ms:.3f} ms  busbw(fp32-equivalent) "
              f"{2 * (world - 1) / world * nbytes / ms / 1e6:.0f} GB/s   max |p2p - nccl| = {err:.3e}  rel L2 {rel:.2e}   "
              f"identical on every rank: {lo.item() == hi.item()}", flush=True)
    ms_b = timed(grp.barrier)
    if rank == 0 and wire == "fp32":
        print(f"one barrier kernel: {ms_b:.3f} ms")
    grp.close()
dist.destroy_process_group()
