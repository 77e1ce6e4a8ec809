"""Symmetric (CUDA-IPC, peer-mapped) buffers for the distributed contrastive loss.

Reference behaviour being replaced (utils/distributed.py:28-58 + modules/losses/contrastive_loss_with_temperature.py
:26-47): two NCCL all_gathers (+ torch.cat) in the forward, two reduce_scatters in the backward.  Here every rank
publishes its bf16 embeddings and its row-LSE vectors in its own symmetric buffer; peers read them in place:
the TMA producer of the similarity / gradient GEMMs loads the remote tiles straight over NVLink, and the GLOBAL
backward needs no gradient traffic at all (engine_loss.contrastive_schedule).  torch.distributed is used ONCE, at
set-up, to exchange the 64-byte IPC handles.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Tuple

import torch
import torch.distributed as dist

from . import _lib
from ._lib import MMBError


class _RawCuda:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can alias it (no copy)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def raw_tensor(ptr: int, nbytes: int, device) -> torch.Tensor:
    return torch.as_tensor(_RawCuda(ptr, nbytes), device=device)


def _align(n: int, a: int = 256) -> int:
    return -(-n // a) * a


class _Slots:
    """Typed views into one rank's symmetric buffer (mine or a mapped peer's)."""

    def __init__(self, raw: torch.Tensor, B: int, E: int):
        off = 0
        self.a, self.b, self.lse_a, self.lse_b, self.w = [], [], [], [], []
        for _ in range(2):
            for lst, nbytes, dt, shape in ((self.a, B * E * 2, torch.bfloat16, (B, E)), (self.b, B * E * 2, torch.bfloat16, (B, E)),
                                           (self.lse_a, B * 4, torch.float32, (B,)), (self.lse_b, B * 4, torch.float32, (B,)),
                                           (self.w, B * 4, torch.float32, (B,))):
                lst.append(raw[off:off + nbytes].view(dt).view(shape))
                off += _align(nbytes)
        self.flags = raw[off:off + 128].view(torch.int32)
        off += 128
        self.nbytes = off

    @staticmethod
    def size(B: int, E: int) -> int:
        return 2 * (2 * _align(B * E * 2) + 3 * _align(B * 4)) + 128


class SymmComm:
    """One per (process, B, E).  Collective constructor: every rank must create it at the same point."""

    _cache: Dict[Tuple[int, int, int], "SymmComm"] = {}

    @classmethod
    def get(cls, B: int, E: int, device) -> "SymmComm":
        key = (B, E, device.index if device.index is not None else torch.cuda.current_device())
        c = cls._cache.get(key)
        if c is None:
            c = cls._cache[key] = SymmComm(B, E, device)
        return c

    def __init__(self, B: int, E: int, device):
        if not (dist.is_available() and dist.is_initialized()):
            raise MMBError("SymmComm needs an initialised torch.distributed process group (handle exchange)")
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.B, self.E, self.device = B, E, device
        L = _lib.lib()
        nbytes = _Slots.size(B, E)
        ptr = ctypes.c_void_p(0)
        _lib.check(L.mmb_symm_alloc(nbytes, ctypes.byref(ptr)), "mmb_symm_alloc")
        self._ptr = ptr.value
        handle = (ctypes.c_ubyte * 64)()
        _lib.check(L.mmb_symm_get_handle(ctypes.c_void_p(self._ptr), handle), "mmb_symm_get_handle")
        allh = [None] * self.world
        dist.all_gather_object(allh, bytes(handle))  # set-up only; backend-agnostic (nccl or gloo)
        self._peer_ptrs: List[int] = []
        self.slots: List[_Slots] = []
        for r in range(self.world):
            if r == self.rank:
                p = self._ptr
            else:
                hb = (ctypes.c_ubyte * 64)(*allh[r])
                pp = ctypes.c_void_p(0)
                _lib.check(L.mmb_symm_open_handle(hb, ctypes.byref(pp)), "mmb_symm_open_handle (is P2P/IPC available?)")
                p = pp.value
            self._peer_ptrs.append(p)
            self.slots.append(_Slots(raw_tensor(p, nbytes, device), B, E))
        self.my = self.slots[self.rank]
        self._flag_table = torch.tensor([s.flags.data_ptr() for s in self.slots], dtype=torch.int64, device=device)
        self._epoch = 0
        self.step = 0
        dist.barrier()  # every rank has mapped every buffer before anybody signals

    def barrier(self) -> None:
        """Cross-GPU barrier on the current stream (release my writes / acquire the peers')."""
        self._epoch += 1
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.lib().mmb_symm_signal_wait(ctypes.c_void_p(self._flag_table.data_ptr()),
                                                   ctypes.c_void_p(self.my.flags.data_ptr()), self.rank, self.world,
                                                   self._epoch, st), "mmb_symm_signal_wait")


class _LocalSlots:
    """world_size == 1: same layout in ordinary device memory, no flags."""

    def __init__(self, B: int, E: int, device):
        raw = torch.zeros(_Slots.size(B, E), dtype=torch.uint8, device=device)
        self.slots = [_Slots(raw, B, E)]
        self.my = self.slots[0]
        self.world, self.rank, self.step = 1, 0, 0

    def barrier(self) -> None:
        return None


_local_cache: Dict[Tuple[int, int, int], _LocalSlots] = {}


def get_comm(B: int, E: int, device, world: int):
    if world > 1:
        return SymmComm.get(B, E, device)
    key = (B, E, device.index if device.index is not None else torch.cuda.current_device())
    c = _local_cache.get(key)
    if c is None:
        c = _local_cache[key] = _LocalSlots(B, E, device)
    return c
