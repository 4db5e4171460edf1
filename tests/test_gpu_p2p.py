"""The gradient-exchange kernels of csrc/p2p.cu on ONE GPU: the "ranks" are buffers of this process and each rank's
kernel runs in turn on the same stream (rank r reads slice r of every buffer and writes slice r of every buffer, so the
sequential launches compute exactly what the concurrent ones do between their barriers).  The cross-process path (CUDA
IPC mapping, the device-side barrier) is exercised by bench.py under torchrun and profiles/tools/allreduce_probe.py."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ptrs(ts):
    return (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("n", [8 * 1024 * 3 + 5, 1 << 20])
def test_two_shot_exchange_fp32(cuda, world, n):
    from mm_s2ut_b200 import _lib

    lib = _lib.load()
    g = torch.Generator(device=cuda).manual_seed(world * 1000 + n % 97)
    bufs = [torch.randn(n, device=cuda, generator=g) for _ in range(world)]
    want = torch.stack(bufs).double().sum(0)
    ref = bufs[0].clone()
    for b in bufs[1:]:
        ref += b                                  # the kernel's summation order: rank 0, 1, 2, ...
    s = torch.cuda.current_stream().cuda_stream
    for r in range(world):
        _lib.check(lib.mm_p2p_allreduce_f32(_ptrs(bufs), world, r, n, s), "mm_p2p_allreduce_f32")
    torch.cuda.synchronize()
    for b in bufs:
        assert torch.equal(b, ref)                # bit-identical on every rank, fixed order
    assert (ref.double() - want).abs().max().item() < 1e-5


@pytest.mark.parametrize("world", [2, 5, 8])
def test_two_shot_exchange_bf16_wire(cuda, world):
    """pack -> exchange on the bf16 staging buffers -> unpack: the sum of the bf16-rounded contributions, accumulated
    in fp32 in rank order and rounded once; identical on every rank; zero tail beyond n."""
    from mm_s2ut_b200 import _lib

    lib = _lib.load()
    n = 8 * world * 777 + 13
    n_pad = (n + 8 * world - 1) // (8 * world) * (8 * world)
    g = torch.Generator(device=cuda).manual_seed(world)
    grads = [torch.randn(n, device=cuda, generator=g) * 10 ** float(i % 3 - 1) for i in range(world)]
    stages = [torch.full((n_pad,), 7.0, dtype=torch.bfloat16, device=cuda) for _ in range(world)]
    s = torch.cuda.current_stream().cuda_stream
    for r in range(world):
        _lib.check(lib.mm_p2p_pack_bf16(grads[r].data_ptr(), stages[r].data_ptr(), n, n_pad, s), "mm_p2p_pack_bf16")
    torch.cuda.synchronize()
    for r in range(world):
        assert torch.equal(stages[r][:n], grads[r].to(torch.bfloat16)) and not stages[r][n:].any()
    acc = torch.zeros(n, device=cuda)
    for r in range(world):
        acc += grads[r].to(torch.bfloat16).float()
    want = acc.to(torch.bfloat16).float()
    for r in range(world):
        _lib.check(lib.mm_p2p_allreduce_bf16(_ptrs(stages), world, r, n_pad, s), "mm_p2p_allreduce_bf16")
    out = [torch.empty(n, device=cuda) for _ in range(world)]
    for r in range(world):
        _lib.check(lib.mm_p2p_unpack_bf16(stages[r].data_ptr(), out[r].data_ptr(), n, s), "mm_p2p_unpack_bf16")
    torch.cuda.synchronize()
    for r in range(world):
        assert torch.equal(out[r], want)
    exact = torch.stack(grads).double().sum(0)
    rel = ((want.double() - exact).norm() / exact.norm()).item()
    assert rel < 6e-3, rel                        # bf16 rounding of the contributions and of the sum
