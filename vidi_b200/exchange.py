"""Peer-memory exchange of the text stream's cross-attention partials (kernels in csrc/xchg.cu).

Replaces the reference's sequence-parallel ``Gather.forward`` (lmm/dattn/sequence_parallel/all_to_all.py:361, split.py:72-93):
every rank owns an *arena* (cudaMalloc, exported to the other ranks of the node through CUDA IPC); per layer each rank reduces
its own key splits to one (O, LSE) partial per stream and stores it into its block of every peer's arena over NVLink, then
publishes a sequence number; the merge kernel of each rank waits on its own flag words.  torch.distributed is used once, at
construction, to swap the 64-byte IPC handles.

Arena layout (bytes):  data fp32 [2 slots][world][cap] | flags uint32 [2 slots][world] | counter uint32 | err int32
Block of one rank inside a slot (floats): stream 0: O [rows, dh] | LSE [rows]; stream 1: O [rows, dh] | LSE [rows]
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import lib as _lib


class _RawCuda:
    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = dict(shape=(nbytes,), typestr="|u1", data=(ptr, False), version=2)


def arena_bytes(world: int, cap: int) -> int:
    return 2 * world * cap * 4 + 2 * world * 4 + 8


class PartialExchange:
    def __init__(self, rank: int, world: int, arenas: Sequence[int], cap: int, device, keep=None, owned: Optional[int] = None,
                 opened: Sequence[int] = ()):
        assert 0 <= rank < world <= 16 and len(arenas) == world
        self.rank, self.world, self.arenas, self.cap, self.device = rank, world, [int(a) for a in arenas], cap, torch.device(device)
        self._keep, self._owned, self._opened = keep, owned, list(opened)
        self.seq = 0
        self.flags_off = 2 * world * cap * 4
        self.counter_ptr = self.arenas[rank] + self.flags_off + 2 * world * 4
        self.err_ptr = self.counter_ptr + 4
        with torch.cuda.device(self.device):
            self._mine = torch.as_tensor(_RawCuda(self.arenas[rank], arena_bytes(world, cap)), device=self.device)
        self._err = self._mine[self.flags_off + 2 * world * 4 + 4:][:4].view(torch.int32)

    # ------------------------------------------------------------------------------------------------
    @staticmethod
    def capacity(max_text_rows: int, dh: int, nstreams: int = 2) -> int:
        return nstreams * max_text_rows * (dh + 1)

    @classmethod
    def local_group(cls, world: int, cap: int, device="cuda") -> List["PartialExchange"]:
        """``world`` exchanges whose arenas all live on ONE device (plain torch buffers): the single-process stand-in the tests use to
        run every rank's text pass in lock step through exactly the kernels and addressing of the multi-process path."""
        bufs = [torch.zeros(arena_bytes(world, cap), device=device, dtype=torch.uint8) for _ in range(world)]
        ptrs = [b.data_ptr() for b in bufs]
        return [cls(r, world, ptrs, cap, device, keep=bufs) for r in range(world)]

    @classmethod
    def connect(cls, rank: int, world: int, group, cap: int, device) -> "PartialExchange":
        """Collective over ``group``: allocate this rank's arena, swap IPC handles, map every peer's arena.  Raises on any rank's
        failure only after all ranks have agreed (so that the caller can fall back to the NCCL exchange on every rank alike)."""
        import torch.distributed as dist
        L = _lib.load()
        device = torch.device(device)
        ok, mine, handle, peers, opened, err = 1, C.c_void_p(), (C.c_ubyte * 64)(), [], [], ""
        with torch.cuda.device(device):
            try:
                _lib.check(L.vidi_p2p_alloc(arena_bytes(world, cap), C.byref(mine), handle), "p2p_alloc")
            except Exception as e:  # noqa: BLE001
                ok, err = 0, repr(e)
            h = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device=device)
            allh = torch.empty(world * 64, dtype=torch.uint8, device=device)
            dist.all_gather_into_tensor(allh, h, group=group)
            allh = allh.cpu().view(world, 64)
            if ok:
                try:
                    for r in range(world):
                        if r == rank:
                            peers.append(mine.value)
                            continue
                        p = C.c_void_p()
                        hb = (C.c_ubyte * 64)(*allh[r].tolist())
                        _lib.check(L.vidi_p2p_open(hb, C.byref(p)), "p2p_open")
                        peers.append(p.value)
                        opened.append(p.value)
                except Exception as e:  # noqa: BLE001
                    ok, err = 0, repr(e)
            flag = torch.tensor([ok], dtype=torch.int32, device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 0:
                for p in opened:
                    L.vidi_p2p_close(p)
                if mine.value:
                    L.vidi_p2p_free(mine)
                raise RuntimeError(f"peer-memory exchange unavailable on some rank (this rank: {err or 'ok'})")
        return cls(rank, world, peers, cap, device, owned=mine.value, opened=opened)

    def close(self):
        L = _lib.load()
        for p in self._opened:
            L.vidi_p2p_close(p)
        self._opened = []
        if self._owned:
            L.vidi_p2p_free(self._owned)
            self._owned = None

    # ------------------------------------------------------------------------------------------------
    def fits(self, nstreams: int, rows: int, dh: int) -> bool:
        return nstreams * rows * (dh + 1) <= self.cap

    def push(self, srcs, rows: int, dh: int, stream: int) -> int:
        """srcs: [(O ptr, LSE ptr, splits)] (this rank's split partials per stream).  Returns the sequence number of this exchange."""
        L = _lib.load()
        assert 1 <= len(srcs) <= 2 and self.fits(len(srcs), rows, dh)
        self.seq += 1
        slot = self.seq & 1
        w = self.world
        base = (C.c_void_p * w)(*[a + slot * w * self.cap * 4 for a in self.arenas])
        flag = (C.c_void_p * w)(*[a + self.flags_off + (slot * w + self.rank) * 4 for a in self.arenas])
        s0 = srcs[0]
        s1 = srcs[1] if len(srcs) > 1 else (None, None, 0)
        _lib.check(L.vidi_xattn_premerge_push(s0[0], s0[1], s0[2], s1[0], s1[1], s1[2], len(srcs), rows, dh, base, flag, w,
                                              self.rank * self.cap, self.seq, self.counter_ptr, stream), "xattn_premerge_push")
        return self.seq

    def merge(self, gates, att: torch.Tensor, out_bf16: torch.Tensor, rows: int, dh: int, stream: int):
        """out = bf16(att + sum_s gate_s * merge over ranks of stream s), after every rank's push of the current sequence number."""
        L = _lib.load()
        slot = self.seq & 1
        w = self.world
        data = self.arenas[self.rank] + slot * w * self.cap * 4
        flags = self.arenas[self.rank] + self.flags_off + slot * w * 4
        a = []
        for s in range(2):
            if s < len(gates):
                o = data + s * rows * (dh + 1) * 4
                a += [o, o + rows * dh * 4, w, 1, self.cap, self.cap, float(gates[s])]
            else:
                a += [None, None, 0, 1, 0, 0, 0.0]
        _lib.check(L.vidi_xattn_merge2_sync(*a, len(gates), att.data_ptr(), rows, dh, out_bf16.data_ptr(), flags, w, self.seq,
                                            self.err_ptr, stream), "xattn_merge2_sync")

    def check(self):
        """Host-side check (synchronises): a peer that never published shows up here instead of as a hang."""
        if int(self._err.item()) != 0:
            raise RuntimeError("vidi_b200: a peer rank never delivered its cross-attention partials (exchange timed out)")
