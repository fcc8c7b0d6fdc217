"""Worker-side replica of a sharded table with background delta exchange (sender-side combining).

``ReplicaCache(table)`` pulls a full local copy of ``table`` (row index == id).  Fused kernels train
against ``replica.table_c`` (a single-shard ``ShardTable`` over local HBM); every ``sync_every`` calls
of :meth:`after_step` a background stream runs ``fps_cache_sync``:

  phase A: push ``replica - base`` to the master shards (one REDG per changed 16-byte chunk),
           ``base <- value read``;
  phase B: ``replica += master - base`` (the other workers' contributions, local REDG),
           ``base <- master``.

The invariant ``replica - base == local updates not yet pushed`` makes the exchange safe against
concurrently running training kernels; nothing is lost, nothing needs a barrier, and a row crosses
NVLink once per exchange instead of once per update.  This is the aggregated form of the reference's
count / timer batching senders (M/common/CombinationLogic.scala) -- see DESIGN.md §2.1.
"""
from __future__ import annotations

import os
import time
from typing import List, Optional

import torch

from ..ops import native
from ..utils.metrics import GLOBAL as METRICS
from .sharded_table import ShardedTable


class ReplicaCache:
    def __init__(self, table: ShardedTable, sync_every: int = 4, sync_interval_ms: Optional[float] = None,
                 require: str = "any", exchange_ctas: Optional[int] = None, overlap_steps: int = 2):
        """Exchange trigger = the reference's combinable conditions (CountLogic / TimerLogic,
        CombinationWorkerSender): ``sync_every`` micro-batches (count), ``sync_interval_ms`` since the
        last exchange (timer), combined with ``require="any"`` (OR, default) or ``"all"`` (AND)."""
        self.table = table
        self.sync_every = max(1, int(sync_every))
        # optional co-scheduling: confine the exchange kernel to `exchange_ctas` CTAs per SM and let the
        # training kernels of the next `overlap_steps` micro-batches leave that many slots free
        # (`reserve()`).  Default 0 = uncoordinated full-size grids: measured faster (N=1, exchange every
        # 4 steps: 0.615 ms/step vs 0.666 / 0.661 with 1 / 2 reserved CTAs) -- a full-width exchange
        # finishes in ~0.2 ms, a confined one slows training for longer than that.
        if exchange_ctas is None:
            exchange_ctas = int(os.environ.get("FPS_EXCHANGE_CTAS", "0"))
        self.exchange_ctas = max(0, int(exchange_ctas))
        self.overlap_steps = int(os.environ.get("FPS_EXCHANGE_OVERLAP_STEPS", overlap_steps))
        self._steps_after_exchange = 1 << 30
        self.sync_interval = None if sync_interval_ms is None else float(sync_interval_ms) / 1000.0
        if require not in ("any", "all"):
            raise ValueError("require must be 'any' or 'all'")
        self.require = require
        self._last_sync = time.monotonic()
        dev = table.cuda_device
        with torch.cuda.device(table.device):
            n_pad = table.rows_per_shard * table.world
            self.cache = torch.empty((n_pad, table.stride), dtype=torch.float32, device=dev)
            table.barrier()
            native.pull_gather(table.table_c, torch.arange(n_pad, device=dev, dtype=torch.int64), self.cache)
            self.base = self.cache.clone()
            self.table_c = native.local_table(self.cache, table.dim)
            self.stream = torch.cuda.Stream(device=dev)
        self._pending: List[torch.cuda.Event] = []
        self._since_sync = 0
        self.exchanges = 0
        table.barrier()

    def reserve(self) -> int:
        """CTA slots per SM the next training kernel should leave free for an exchange in flight."""
        return self.exchange_ctas if self._steps_after_exchange < self.overlap_steps else 0

    def after_step(self) -> None:
        self._since_sync += 1
        self._steps_after_exchange += 1
        count_hit = self._since_sync >= self.sync_every
        if self.sync_interval is None:
            fire = count_hit
        else:
            timer_hit = time.monotonic() - self._last_sync >= self.sync_interval
            fire = (count_hit or timer_hit) if self.require == "any" else (count_hit and timer_hit)
        if fire:
            self.exchange()

    def exchange(self) -> None:
        """Start one delta exchange on the background stream (overlaps later training kernels)."""
        dev = self.table.cuda_device
        cur = torch.cuda.current_stream(dev)
        ev = torch.cuda.Event()
        ev.record(cur)
        self.stream.wait_event(ev)                    # include everything trained so far
        with torch.cuda.stream(self.stream):
            native.cache_sync(self.table.table_c, self.cache, self.base, self.exchange_ctas)
            done = torch.cuda.Event()
            done.record(self.stream)
        self._pending.append(done)
        if len(self._pending) > 2:                    # at most two exchanges outstanding
            cur.wait_event(self._pending.pop(0))
        self._since_sync = 0
        self._steps_after_exchange = 0
        self._last_sync = time.monotonic()
        self.exchanges += 1
        METRICS.inc("replica_exchanges")

    def flush(self) -> None:
        """Push every pending local delta to the masters and make the current stream wait for it."""
        if self._since_sync > 0:
            self.exchange()
        cur = torch.cuda.current_stream(self.table.cuda_device)
        for ev in self._pending:
            cur.wait_event(ev)
        self._pending = []
