"""Peer-memory gradient exchange (``csrc/p2p.cu``): every rank maps the other ranks' flat gradient buffers (and a small
flag array) through CUDA IPC once; ``all_reduce`` is then three plain kernel launches -- barrier, two-shot in-place
exchange, barrier -- on the current stream, so it can be captured into a CUDA graph and run on a side stream beside the
backward pass.

The reference trains data-parallel through fairseq's DDP wrapper (scripts/textless/1_train.sh:105-125,
``--distributed-world-size``), i.e. an NCCL all-reduce of the gradients.  This is the same exchange for this path's flat
fp32 gradient buffer, written against NVLink peer memory: one process per GPU, one node.

The recipe trains with ``--fp16`` (1_train.sh:125), so fairseq's exchange moves 16-bit gradients.  By default the peer
exchange does the same: the fp32 gradients are rounded into a bf16 staging buffer (the buffer the peers map), summed in
fp32 in rank order with one rounding of the sum, and widened back -- half the NVLink bytes, every rank ends with the
same bits.  ``MM_P2P_GRAD_DTYPE=fp32`` exchanges the fp32 buffer itself.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch

from . import _lib


def _exchange_handles(t: torch.Tensor):
    """CUDA IPC handle + offset of ``t`` gathered over the ranks -> this rank's view of every rank's tensor (pointers)."""
    import torch.distributed as dist

    lib = _lib.load()
    handle = (C.c_uint8 * 64)()
    off = C.c_int64(0)
    _lib.check(lib.mm_ipc_get_handle(t.data_ptr(), handle, C.byref(off)), "mm_ipc_get_handle")
    world, rank = dist.get_world_size(), dist.get_rank()
    everyone: List = [None] * world
    dist.all_gather_object(everyone, (bytes(handle), int(off.value), int(t.numel())))
    ptrs, mapped = [], []
    for r, (h, o, n) in enumerate(everyone):
        if n != t.numel():
            raise ValueError(f"peer mapping: rank {r} holds {n} elements, this rank {t.numel()}")
        if r == rank:
            ptrs.append(t.data_ptr())
            continue
        base = C.c_void_p()
        buf = (C.c_uint8 * 64).from_buffer_copy(h)
        _lib.check(lib.mm_ipc_open_handle(buf, C.byref(base)), "mm_ipc_open_handle")
        mapped.append(base.value)
        ptrs.append(base.value + o)
    return ptrs, mapped


class PeerGroup:
    """Maps ``tensor`` (same shape on every rank of the default process group) into every rank's address space."""

    def __init__(self, tensor: torch.Tensor, exchange_dtype: Optional[str] = None):
        import os

        import torch.distributed as dist

        assert tensor.is_cuda and tensor.dtype == torch.float32 and tensor.is_contiguous()
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        if not 2 <= self.world <= 8:
            raise ValueError("PeerGroup: 2 .. 8 ranks of one node")
        self.grads = tensor
        self.exchange_dtype = exchange_dtype or os.environ.get("MM_P2P_GRAD_DTYPE", "bf16")
        if self.exchange_dtype not in ("bf16", "fp32"):
            raise ValueError("MM_P2P_GRAD_DTYPE: bf16 or fp32")
        if self.exchange_dtype == "bf16":      # the peers map the 16-bit staging copy, not the gradients themselves
            n_pad = (tensor.numel() + 8 * self.world - 1) // (8 * self.world) * (8 * self.world)
            self.stage = torch.zeros(n_pad, dtype=torch.bfloat16, device=tensor.device)
            tensor = self.stage
        self.align = 8 if self.exchange_dtype == "bf16" else 4
        self.tensor = tensor
        # flag array (slot p written by rank p) + the local epoch counter: one dedicated allocation per rank
        self._sync = torch.zeros(64, dtype=torch.int32, device=tensor.device)
        torch.cuda.synchronize(tensor.device)
        self._ptrs, m1 = _exchange_handles(tensor)
        flag_ptrs, m2 = _exchange_handles(self._sync)
        self._mapped = m1 + m2
        self._flags = (C.c_void_p * self.world)(*flag_ptrs)
        self._epoch = self._sync.data_ptr() + 32 * 4          # elements 32.. of the local array: never written by peers
        dist.barrier()          # every rank has zeroed and mapped its flags before the first device-side barrier

    def barrier(self) -> None:
        """Device-side barrier kernel on the current stream (capturable): see ``mm_p2p_barrier``."""
        from . import kernels as K

        lib = _lib.load()
        with K._Launch("p2p_barrier"):
            _lib.check(lib.mm_p2p_barrier(self._flags, self._epoch, self.world, self.rank, K._stream()), "mm_p2p_barrier")

    def all_reduce(self, lo: int = 0, hi: Optional[int] = None) -> None:
        """grads[lo:hi] <- sum over ranks, in place on every rank (fixed summation order: bit-identical everywhere; in
        bf16 mode the contributions and the sum are rounded to bf16).  lo and hi must be multiples of ``align``
        elements (16-byte vector accesses) unless hi is the end of the tensor."""
        from . import kernels as K

        n = self.grads.numel()
        hi = n if hi is None else hi
        if hi <= lo:
            return
        assert lo % self.align == 0 and (hi % self.align == 0 or hi == n), (lo, hi)
        lib = _lib.load()
        if self.exchange_dtype == "bf16":
            hi_pad = self.stage.numel() if hi == n else hi          # the zeroed tail rides along
            g, st = self.grads.data_ptr() + 4 * lo, self.stage.data_ptr() + 2 * lo
            ptrs = (C.c_void_p * self.world)(*[p + 2 * lo for p in self._ptrs])
            with K._Launch("p2p_pack", 6.0 * (hi - lo)):
                _lib.check(lib.mm_p2p_pack_bf16(g, st, hi - lo, hi_pad - lo, K._stream()), "mm_p2p_pack_bf16")
            self.barrier()          # every rank's staging copy of [lo, hi) is complete
            with K._Launch("p2p_allreduce", 4.0 * (hi_pad - lo) * (self.world - 1) / self.world):
                _lib.check(lib.mm_p2p_allreduce_bf16(ptrs, self.world, self.rank, hi_pad - lo, K._stream()),
                           "mm_p2p_allreduce_bf16")
            self.barrier()          # every rank's stores into this rank's staging buffer have landed
            with K._Launch("p2p_unpack", 6.0 * (hi - lo)):
                _lib.check(lib.mm_p2p_unpack_bf16(st, g, hi - lo, K._stream()), "mm_p2p_unpack_bf16")
            return
        ptrs = (C.c_void_p * self.world)(*[p + 4 * lo for p in self._ptrs])
        self.barrier()              # every rank's gradients in [lo, hi) are complete
        with K._Launch("p2p_allreduce", 8.0 * (hi - lo) * (self.world - 1) / self.world):
            _lib.check(lib.mm_p2p_allreduce_f32(ptrs, self.world, self.rank, hi - lo, K._stream()), "mm_p2p_allreduce_f32")
        self.barrier()              # every rank's stores into this rank's buffer have landed

    def close(self) -> None:
        lib = _lib.load()
        for b in self._mapped:
            lib.mm_ipc_close_handle(b)
        self._mapped = []


_groups: Dict[int, PeerGroup] = {}


def peer_group(flat: torch.Tensor) -> Optional[PeerGroup]:
    """The PeerGroup of ``flat`` (created on first use: a collective call -- every rank must get here), or None when
    this process group cannot use peer memory (not NCCL / not CUDA / more than 8 ranks / MM_P2P_ALLREDUCE=0)."""
    import os

    import torch.distributed as dist

    if os.environ.get("MM_P2P_ALLREDUCE", "1") == "0":
        return None
    if not (flat.is_cuda and flat.dtype == torch.float32 and dist.is_available() and dist.is_initialized()
            and dist.get_backend() == "nccl" and 2 <= dist.get_world_size() <= 8):
        return None
    g = _groups.get(flat.data_ptr())
    if g is None or g.grads.numel() != flat.numel():
        g = _groups[flat.data_ptr()] = PeerGroup(flat)
    return g


def peer_all_reduce(flat: torch.Tensor) -> Optional[int]:
    """All-reduce ``flat`` over peer memory if this process group can; returns the world size, or None when the caller
    should use the library collective instead."""
    g = peer_group(flat)
    if g is None:
        return None
    g.all_reduce()
    return g.world
