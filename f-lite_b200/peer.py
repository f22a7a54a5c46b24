"""NVLink peer-memory plumbing for the fused compute+exchange kernels (Ulysses sequence parallelism).

One process per GPU; every rank of a process group on ONE node allocates the same-sized buffer
(``flite_p2p_alloc``), publishes its CUDA IPC handle through ``torch.distributed`` and maps the peers' buffers
(``flite_ipc_open``).  The kernels then store straight into the destination rank's buffer
(``flite_gemm_qkv_p2p`` / ``flite_attention_varlen_p2p``) and completion is published with stream-ordered flag
kernels (``flite_p2p_signal`` / ``flite_p2p_wait``) -- no NCCL call on the data path.

The reference has no multi-GPU inference; this replaces the two ``all_to_all_single`` calls of the
NCCL Ulysses path in ``model.py``.
"""
from __future__ import annotations

import ctypes
from typing import List

import torch
import torch.distributed as dist

from . import _lib

FLAG_BYTES = 256          # 8 x uint32 flag slots, padded to keep the data region 256-byte aligned


class _CudaArray:
    """Minimal ``__cuda_array_interface__`` holder so that torch can view raw device memory."""

    def __init__(self, ptr: int, n_i16: int):
        self.__cuda_array_interface__ = {"shape": (n_i16,), "typestr": "<i2", "data": (ptr, False), "version": 3}


class SymmetricBuffer:
    """``nbytes`` of device memory on every rank of ``group`` + 8 flag slots, mapped into every peer.

    ``ptrs[g]`` / ``flag_ptrs[g]`` are the addresses of rank g's data / flag region valid in THIS process.
    Construction and ``close()`` are collective over ``group``."""

    def __init__(self, group, nbytes: int, device: torch.device):
        lib = _lib.load()
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > 8:
            raise _lib.FliteError("peer-memory exchange supports at most 8 ranks (one NVSwitch box)")
        self.nbytes = (int(nbytes) + 255) // 256 * 256
        self.device = device
        base = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.flite_p2p_alloc(FLAG_BYTES + self.nbytes, ctypes.byref(base)), "p2p_alloc")
            handle = ctypes.create_string_buffer(64)
            _lib.check(lib.flite_ipc_get_handle(base, handle), "ipc_get_handle")
            handles: List[bytes] = [b""] * self.world
            dist.all_gather_object(handles, bytes(handle.raw), group=group)
            self._base = base.value
            self._opened = []
            bases = []
            for g in range(self.world):
                if g == self.rank:
                    bases.append(self._base)
                    continue
                p = ctypes.c_void_p()
                _lib.check(lib.flite_ipc_open(ctypes.create_string_buffer(handles[g], 64), ctypes.byref(p)),
                           f"ipc_open(rank {g})")
                self._opened.append(p.value)
                bases.append(p.value)
        self.flag_ptrs = (ctypes.c_void_p * 8)(*(bases + [None] * (8 - self.world)))
        self.data_bases = [b + FLAG_BYTES for b in bases]
        self.local = torch.as_tensor(_CudaArray(self._base + FLAG_BYTES, self.nbytes // 2), device=device).view(
            torch.bfloat16)
        self.epoch = 0
        dist.barrier(group=group)      # every peer has mapped this rank's buffer before anyone stores into it

    def table(self, byte_offset: int):
        """Host array of the peers' pointers to the sub-buffer at ``byte_offset`` (same offset on every rank)."""
        return (ctypes.c_void_p * 8)(*([b + byte_offset for b in self.data_bases] + [None] * (8 - self.world)))

    def exchange_done(self, stream: int) -> None:
        """Stream-ordered all-to-all completion: publish "my stores up to here are done" to every peer, then block
        the stream until every peer has published the same epoch."""
        from . import ops
        lib = _lib.load()
        self.epoch = (self.epoch + 1) & 0xFFFFFFFF
        tr = ops.TRACE
        if tr is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        _lib.check(lib.flite_p2p_signal(self.flag_ptrs, self.world, self.rank, self.epoch, stream), "p2p_signal")
        _lib.check(lib.flite_p2p_wait(self._base, self.world, self.epoch, stream), "p2p_wait")
        ops.LAUNCHES[0] += 2
        if tr is not None:
            e1.record()
            tr.append(("p2p signal+wait", e0, e1))

    def close(self) -> None:
        if self._base is None:
            return
        lib = _lib.load()
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        self.local = None
        for p in self._opened:
            lib.flite_ipc_close(p)
        self._opened = []
        dist.barrier(group=self.group)
        lib.flite_p2p_free(self._base)
        self._base = None
