"""Peer-memory gradient exchange (``csrc/p2p.cu``): every rank maps the other ranks' flat gradient buffers through
CUDA IPC once; ``all_reduce`` then runs the two-shot in-place kernel between two stream-ordered barriers.

The reference trains data-parallel through fairseq's DDP wrapper (scripts/textless/1_train.sh:105-125,
``--distributed-world-size``), i.e. an NCCL all-reduce of the gradients.  This is the same exchange for this path's flat
fp32 gradient buffer, written against NVLink peer memory: one process per GPU, one node.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch

from . import _lib


class PeerGroup:
    """Maps ``tensor`` (same shape on every rank of the default process group) into every rank's address space."""

    def __init__(self, tensor: torch.Tensor):
        import torch.distributed as dist

        assert tensor.is_cuda and tensor.dtype == torch.float32 and tensor.is_contiguous()
        self.dist = dist
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        if not 2 <= self.world <= 8:
            raise ValueError("PeerGroup: 2 .. 8 ranks of one node")
        self.tensor = tensor
        lib = _lib.load()
        handle = (C.c_uint8 * 64)()
        off = C.c_int64(0)
        _lib.check(lib.mm_ipc_get_handle(tensor.data_ptr(), handle, C.byref(off)), "mm_ipc_get_handle")
        mine = (bytes(handle), int(off.value), int(tensor.numel()), torch.cuda.current_device())
        everyone: List = [None] * self.world
        dist.all_gather_object(everyone, mine)
        self._mapped: List[int] = []
        ptrs = []
        for r, (h, o, n, _dev) in enumerate(everyone):
            if n != tensor.numel():
                raise ValueError(f"PeerGroup: rank {r} holds {n} elements, this rank {tensor.numel()}")
            if r == self.rank:
                ptrs.append(tensor.data_ptr())
                continue
            base = C.c_void_p()
            buf = (C.c_uint8 * 64).from_buffer_copy(h)
            _lib.check(lib.mm_ipc_open_handle(buf, C.byref(base)), "mm_ipc_open_handle")
            self._mapped.append(base.value)
            ptrs.append(base.value + o)
        self._ptrs = (C.c_void_p * self.world)(*ptrs)
        self._flag = torch.zeros(1, device=tensor.device)

    def barrier(self) -> None:
        """Stream-ordered: later work of this stream starts after every rank's earlier work (a 4-byte NCCL all-reduce)."""
        self.dist.all_reduce(self._flag)

    def all_reduce(self) -> None:
        """tensor <- sum over ranks, in place on every rank (fixed summation order: bit-identical everywhere)."""
        from . import kernels as K

        lib = _lib.load()
        self.barrier()              # every rank's gradients are complete
        with K._Launch("p2p_allreduce", 8.0 * self.tensor.numel() * (self.world - 1) / self.world):
            _lib.check(lib.mm_p2p_allreduce_f32(self._ptrs, self.world, self.rank, self.tensor.numel(), K._stream()),
                       "mm_p2p_allreduce_f32")
        self.barrier()              # every rank's stores into this rank's buffer have landed

    def close(self) -> None:
        lib = _lib.load()
        for b in self._mapped:
            lib.mm_ipc_close_handle(b)
        self._mapped = []


_groups: Dict[int, PeerGroup] = {}


def peer_all_reduce(flat: torch.Tensor) -> Optional[int]:
    """All-reduce ``flat`` over peer memory if this process group can (NCCL backend, CUDA tensor, <= 8 ranks); returns
    the world size, or None when the caller should use the library collective instead."""
    import torch.distributed as dist

    if not (flat.is_cuda and dist.is_available() and dist.is_initialized() and dist.get_backend() == "nccl"):
        return None
    world = dist.get_world_size()
    if not 2 <= world <= 8:
        return None
    key = flat.data_ptr()
    g = _groups.get(key)
    if g is None or g.tensor.numel() != flat.numel():
        g = _groups[key] = PeerGroup(flat)
    g.all_reduce()
    return world
