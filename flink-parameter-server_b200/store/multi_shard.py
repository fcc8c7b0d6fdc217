"""``psParallelism`` logical PS shards inside ONE process (single-process device backend).

:class:`ShardedTable` places one shard on every rank of a multi-process job.  ``MultiShardTable`` is its
single-process sibling: ``num_shards`` shards (any number, e.g. the ``psParallelism = 3`` of the
reference's stack tests) are placed round-robin on the GPUs the process can see, with peer access
enabled between them, so the same one-sided kernels address every shard through the same ``ShardTable``
pointer table.  Partitioning: ``hash`` / ``range`` (computed in the kernels) or ``lut`` -- a device
lookup table ``id -> (owner << 40 | slot)`` filled from ANY user partitioner function
(FPS:343 ``paramPartitioner``), one copy per GPU.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch

from ..ops import native
from .sharded_table import ShardedTable


class MultiShardTable(ShardedTable):
    def __init__(self, num_ids: int, dim: int, num_shards: int, *, partition: str = "hash",
                 partitioner: Optional[Callable[[int], int]] = None,
                 devices: Optional[Sequence[int]] = None, init: str = "uniform",
                 init_range=(-0.01, 0.01), seed: int = 0, track_touched: bool = False):
        if not 1 <= num_shards <= native.FPS_MAX_SHARDS:
            raise ValueError(f"num_shards must be in [1, {native.FPS_MAX_SHARDS}]")
        self.dim = int(dim)
        self.stride = (self.dim + 3) // 4 * 4
        self.num_ids = int(num_ids)
        self.group = None
        self.world, self.rank = 1, 0
        self.n_shards = int(num_shards)
        self.owns_shard = True
        if devices is None:
            devices = list(range(torch.cuda.device_count()))
        self.devices = [int(d) for d in devices][: self.n_shards] or [torch.cuda.current_device()]
        self.device = self.devices[0]
        for a in self.devices:
            for b in self.devices:
                if a != b:
                    native.enable_peer(a, b)
        if partitioner is not None:
            partition = "lut"
        self.partition = partition
        self.mode = {"hash": native.PART_HASH, "range": native.PART_RANGE, "lut": native.PART_LUT}[partition]
        self.seed = int(seed)
        self.track_touched = bool(track_touched)

        # ---- id -> (owner, slot) ------------------------------------------------------------------
        self._lut_host = None
        if self.mode == native.PART_LUT:
            owners = torch.tensor([int(partitioner(i)) % self.n_shards for i in range(self.num_ids)],
                                  dtype=torch.int64)
            slots = torch.zeros(self.num_ids, dtype=torch.int64)
            counts = []
            for o in range(self.n_shards):
                sel = (owners == o).nonzero(as_tuple=True)[0]
                slots[sel] = torch.arange(sel.numel())
                counts.append(int(sel.numel()))
            self.rows_per_shard = max(1, max(counts))
            self._lut_host = (owners << native.LUT_OWNER_SHIFT) | slots
            self._owner_ids = [(owners == o).nonzero(as_tuple=True)[0] for o in range(self.n_shards)]
        else:
            self.rows_per_shard = -(-self.num_ids // self.n_shards)
        self.div = self.rows_per_shard

        # ---- shard blocks, round-robin over the GPUs ---------------------------------------------
        row_bytes = self.rows_per_shard * self.stride * 4
        self._bitmap_words = (self.rows_per_shard + 31) // 32
        self._bitmap_off = (row_bytes + 255) // 256 * 256
        total = self._bitmap_off + (self._bitmap_words * 4 if track_touched else 0)
        self._ptrs: List[int] = []
        self.shard_device: List[int] = []
        self.shards: List[torch.Tensor] = []
        self._touched_views: List[Optional[torch.Tensor]] = []
        for s in range(self.n_shards):
            d = self.devices[s % len(self.devices)]
            with torch.cuda.device(d):
                ptr = native.heap_alloc(total)
            self._ptrs.append(ptr)
            self.shard_device.append(d)
            self.shards.append(native.tensor_from_ptr(ptr, (self.rows_per_shard, self.stride), torch.float32, d))
            self._touched_views.append(
                native.tensor_from_ptr(ptr + self._bitmap_off, (self._bitmap_words,), torch.int32, d)
                if track_touched else None)
        self.local = self.shards[0]
        self.touched = self._touched_views[0]
        self._tables = {}
        self._luts = {}
        self.table_c = self.table_for(self.device)
        self.nan_flag = torch.zeros(1, dtype=torch.int32, device=self.cuda_device)
        if init == "uniform":
            if self.mode == native.PART_LUT:
                raise ValueError("Philox init-by-id needs a computed partition; use init='zeros' with a LUT")
            self.init_uniform(*init_range)
        elif init != "zeros":
            raise ValueError(f"unknown init {init!r}")
        self.barrier()

    def table_for(self, device: int) -> native.ShardTableC:
        """The pointer table as seen from ``device`` (the LUT copy is local to the reader)."""
        device = int(device)
        if device not in self._tables:
            tc = native.ShardTableC()
            for s in range(self.n_shards):
                tc.base[s] = self._ptrs[s]
                tc.touched[s] = (self._ptrs[s] + self._bitmap_off) if self.track_touched else None
            tc.rows_per_shard = self.rows_per_shard
            tc.div = self.div
            tc.num_shards = self.n_shards
            tc.dim, tc.stride, tc.mode = self.dim, self.stride, self.mode
            tc.shard_shift = native.log2_or_neg(self.n_shards)
            if self._lut_host is not None:
                self._luts[device] = self._lut_host.to(torch.device("cuda", device))
                tc.lut = self._luts[device].data_ptr()
            self._tables[device] = tc
        return self._tables[device]

    def init_uniform(self, lo: float, hi: float) -> None:
        for s, t in enumerate(self.shards):
            with torch.cuda.device(self.shard_device[s]):
                native.init_rows(t, self.dim, s, self.n_shards, self.mode, self.div, self.seed, lo, hi)

    def shard_ids(self, s: int) -> torch.Tensor:
        dev = torch.device("cuda", self.shard_device[s])
        if self.mode == native.PART_LUT:
            return self._owner_ids[s].to(dev)
        slots = torch.arange(self.rows_per_shard, device=dev, dtype=torch.int64)
        return slots * self.n_shards + s if self.mode == native.PART_HASH else slots + s * self.div

    def local_ids(self) -> torch.Tensor:
        return self.shard_ids(0)

    def dump_local(self, only_touched: Optional[bool] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """(ids, values) of EVERY shard (they are all local to this process)."""
        self.barrier()
        if only_touched is None:
            only_touched = self.track_touched
        ids_out, vals_out = [], []
        for s, t in enumerate(self.shards):
            ids = self.shard_ids(s)
            n = ids.numel()
            valid = ids < self.num_ids
            if only_touched and self._touched_views[s] is not None:
                slots = torch.arange(n, device=ids.device)
                words = self._touched_views[s]
                valid = valid & (((words[slots >> 5] >> (slots & 31)) & 1) != 0)
            sel = valid.nonzero(as_tuple=True)[0]
            ids_out.append(ids[sel].to(self.cuda_device))
            vals_out.append(t[:n][sel, : self.dim].to(self.cuda_device))
        return torch.cat(ids_out), torch.cat(vals_out)

    def barrier(self) -> None:
        for d in set(self.shard_device):
            torch.cuda.synchronize(d)

    def close(self) -> None:
        self.barrier()
        ptrs, self._ptrs = self._ptrs, []
        self.shards, self._touched_views, self.local, self.touched = [], [], None, None
        for p, d in zip(ptrs, self.shard_device):
            with torch.cuda.device(d):
                try:
                    native.heap_free(p)
                except RuntimeError:
                    pass

    def __del__(self):  # best effort
        try:
            if self._ptrs:
                self.close()
        except Exception:
            pass
