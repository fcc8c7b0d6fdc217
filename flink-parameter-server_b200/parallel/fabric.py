"""Symmetric-heap fabric over NVLink/NVSwitch.

Each PS rank owns one device allocation; all ranks map all allocations, so a kernel running on
any GPU can address any shard by pointer.  This is the B200-native replacement of the
reference's ``partitionCustom`` + Flink network stack (FPS:416-420, 455-463) and of its
iteration feedback edge (FPS:477-480): a pull is a peer load, a push is a peer reduction.

Bootstrap uses ``torch.distributed`` (NCCL or gloo) only to exchange 64-byte CUDA IPC handles.
A fallback to ``torch.distributed._symmetric_memory`` (CUDA VMM + fd passing) is used when
legacy CUDA IPC is refused by the container.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.distributed as dist

from ..ops import native


def _dist_ready() -> bool:
    return dist.is_available() and dist.is_initialized()


class SymmetricHeap:
    """A same-size device allocation on every rank of ``group`` with all peers mapped."""

    def __init__(self, nbytes: int, group=None, device: Optional[int] = None,
                 mode: Optional[str] = None):
        self.nbytes = int((nbytes + 255) // 256 * 256)
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.group = group
        self.world = dist.get_world_size(group) if _dist_ready() else 1
        self.rank = dist.get_rank(group) if _dist_ready() else 0
        self.mode = mode or os.environ.get("FPS_FABRIC", "ipc")
        self._opened: List[int] = []
        self._symm_keepalive = None
        self._owned = False
        self.peer_ptrs: List[int] = []
        with torch.cuda.device(self.device):
            if self.world == 1:
                self.local_ptr = native.heap_alloc(self.nbytes)
                self._owned = True
                self.peer_ptrs = [self.local_ptr]
            elif self.mode == "symm":
                self._init_symm()
            else:
                try:
                    self._init_ipc()
                except RuntimeError as e:  # container refuses legacy IPC -> VMM path
                    if os.environ.get("FPS_FABRIC") == "ipc":
                        raise
                    self._cleanup_ipc()
                    self.mode = "symm"
                    self._init_symm()

    # -- CUDA IPC -------------------------------------------------------------------------
    def _init_ipc(self) -> None:
        self.local_ptr = native.heap_alloc(self.nbytes)
        self._owned = True
        handle = native.ipc_get_handle(self.local_ptr)
        handles: List[Optional[bytes]] = [None] * self.world
        dist.all_gather_object(handles, handle, group=self.group)
        ok = 1
        ptrs = []
        err = None
        for r, h in enumerate(handles):
            if r == self.rank:
                ptrs.append(self.local_ptr)
                continue
            try:
                p = native.ipc_open_handle(h)
                self._opened.append(p)
                ptrs.append(p)
            except RuntimeError as e:
                ok = 0
                err = e
                break
        flags = [None] * self.world
        dist.all_gather_object(flags, ok, group=self.group)
        if not all(flags):
            raise RuntimeError(f"CUDA IPC mapping failed on some rank: {err}")
        self.peer_ptrs = ptrs

    def _cleanup_ipc(self) -> None:
        for p in self._opened:
            try:
                native.ipc_close(p)
            except RuntimeError:
                pass
        self._opened = []
        if self._owned:
            try:
                native.heap_free(self.local_ptr)
            except RuntimeError:
                pass
            self._owned = False

    # -- torch symmetric memory (CUDA VMM) ------------------------------------------------
    def _init_symm(self) -> None:
        import torch.distributed._symmetric_memory as symm

        grp = self.group if self.group is not None else dist.group.WORLD
        t = symm.empty(self.nbytes, dtype=torch.uint8, device=torch.device("cuda", self.device))
        hdl = symm.rendezvous(t, grp)
        t.zero_()
        self._symm_keepalive = (t, hdl)
        self.peer_ptrs = [int(p) for p in hdl.buffer_ptrs]
        self.local_ptr = self.peer_ptrs[self.rank]
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)

    # -- views ----------------------------------------------------------------------------
    def local_tensor(self, shape, dtype: torch.dtype = torch.float32, offset: int = 0) -> torch.Tensor:
        return native.tensor_from_ptr(self.local_ptr + offset, shape, dtype, self.device)

    def peer_tensor(self, rank: int, shape, dtype: torch.dtype = torch.float32,
                    offset: int = 0) -> torch.Tensor:
        """Zero-copy view of a *peer's* allocation (loads/stores go over NVLink)."""
        return native.tensor_from_ptr(self.peer_ptrs[rank] + offset, shape, dtype, self.device)

    def barrier(self) -> None:
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(group=self.group)

    def close(self) -> None:
        if self.world > 1:
            try:
                self.barrier()
            except Exception:
                pass
        self._cleanup_ipc()
        self._symm_keepalive = None

    def __del__(self):  # best effort
        try:
            self._cleanup_ipc()
        except Exception:
            pass
