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
dist.destroy_process_group()
