"""Worker-side replica of a sharded table with a device-driven delta exchange (sender-side combining).

``ReplicaCache(table)`` keeps a full local copy of ``table`` laid out **owner-major**: segment ``o``
holds the rows of PS shard ``o`` in slot order, so segment ``o`` is the per-destination send buffer of
the reference's batching senders (M/common/CombinationLogic.scala:12-33,
M/client/sender/CombinationWorkerSender.scala:9-36).  Fused kernels train against
``replica.table_c`` (the master's ``id -> (owner, slot)`` map with every "shard" in local HBM).

After every micro-batch two kernels run on a high-priority side stream (``ops/csrc/fps_replica.cu``):

* ``fps_flush_policy`` -- the **device-side** CountLogic / TimerLogic (CountLogic.scala:5-29,
  TimerLogic.scala:6-51): per destination, "messages buffered >= count" and / or "``globaltimer``
  deadline passed", combined with OR (``require="any"``) or AND (``"all"``).  The message counters
  are fed on the device by the bucket histogram of the training step; no host clock is involved.
* ``fps_replica_exchange`` -- for the flagged destinations: push ``replica - base`` (REDG over
  NVLink), fold ``master - base`` (the other workers' pushes) into the replica, ``base <- master +
  pushed delta``.  A few CTAs stream the segments with TMA bulk copies through a shared-memory ring, so
  the kernel runs *next to* the HBM-bound training kernel (which leaves those CTA slots free).

The invariant ``replica - base == local updates not yet pushed`` holds element-wise whatever the
training kernels do concurrently, so nothing is lost and nothing needs a barrier; a row crosses NVLink
once per flush per direction instead of once per update.  Destinations are staggered across steps and
ranks (``(o - rank) % sync_every``) so the links stay evenly busy.  See DESIGN.md §2.1.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch

from ..ops import native
from ..utils.metrics import GLOBAL as METRICS
from .sharded_table import ShardedTable


class ReplicaCache:
    def __init__(self, table: ShardedTable, sync_every: int = 4, sync_interval_ms: Optional[float] = None,
                 require: str = "any", flush_count: Optional[int] = None,
                 exchange_ctas: Optional[int] = None, stages: Optional[int] = None,
                 stagger: bool = True, max_outstanding: int = 2, own_inplace: Optional[bool] = None):
        """Flush trigger = the reference's combinable conditions: ``flush_count`` messages buffered for
        a destination (CountLogic; default: what ``sync_every`` micro-batches send to one destination),
        ``sync_interval_ms`` since the destination's last flush (TimerLogic, device ``globaltimer``),
        combined with ``require="any"`` (OR) or ``"all"`` (AND)."""
        if require not in ("any", "all"):
            raise ValueError("require must be 'any' or 'all'")
        self.table = table
        self.sync_every = max(1, int(sync_every))
        self.flush_count = None if flush_count is None else int(flush_count)
        self.interval_ns = 0 if sync_interval_ms is None else int(float(sync_interval_ms) * 1e6)
        self.require_all = require == "all"
        self.stagger = bool(stagger)
        # sync_every mode (no explicit count / timer): every micro-batch flushes ONE of `sync_every`
        # rotating slices of every destination buffer -- uniform NVLink / HBM load per step, a row is
        # exchanged once per `sync_every` micro-batches, ranks are staggered so that a peer's delta is
        # seen after ~sync_every/2 steps on average.  An explicit flush_count / sync_interval_ms flushes
        # whole destination buffers when the device-side policy fires.
        self.sliced = flush_count is None and sync_interval_ms is None and self.sync_every > 1
        self.n_ctas = int(os.environ.get("FPS_EXCHANGE_CTAS", 64 if exchange_ctas is None else exchange_ctas))
        self.stages = int(os.environ.get("FPS_EXCHANGE_STAGES", 4 if stages is None else stages))
        self.sequential = os.environ.get("FPS_EXCHANGE_SEQUENTIAL", "0") == "1"
        self.max_outstanding = max(1, int(max_outstanding))
        dev = table.cuda_device
        self.world, self.rank, self.rps = table.n_shards, table.rank, table.rows_per_shard
        with torch.cuda.device(table.device):
            n_rows = self.rps * self.world
            self.cache = torch.empty((n_rows, table.stride), dtype=torch.float32, device=dev)
            table.barrier()
            # segment o, slot s  <-  master row of the id that lives at (owner o, slot s)
            native.pull_gather(table.table_c, self._segment_ids(), self.cache)
            self.base = self.cache.clone()
            # the worker's own shard is trained in place (no replica of it, nothing to exchange)
            if own_inplace is None:
                own_inplace = os.environ.get("FPS_REPLICA_OWN_INPLACE", "1") == "1"
            self.own = table.rank if (table.owns_shard and own_inplace) else None
            self.skip_mask = 0 if self.own is None else (1 << self.own)
            self.table_c = native.segment_table(self.cache, table.table_c, alias=self.own)
            self.state = torch.zeros(native.FLUSH_STATE_WORDS, dtype=torch.int64, device=dev)
            self.stream = torch.cuda.Stream(device=dev, priority=-1)
        self._pending_events: List[torch.cuda.Event] = []
        self._timing = os.environ.get("FPS_EXCHANGE_TIMING", "0") == "1"
        self._timed: List[tuple] = []
        self._armed = False
        self._since_flush = 0
        self.steps = 0
        table.barrier()

    # -- layout ---------------------------------------------------------------------------------
    def _segment_ids(self) -> torch.Tensor:
        dev = self.table.cuda_device
        slots = torch.arange(self.rps, device=dev, dtype=torch.int64)
        owners = torch.arange(self.world, device=dev, dtype=torch.int64)
        if self.table.mode == native.PART_HASH:
            ids = slots[None, :] * self.world + owners[:, None]
        else:
            ids = owners[:, None] * self.table.div + slots[None, :]
        return ids.reshape(-1).contiguous()

    def row_index(self, ids: torch.Tensor) -> torch.Tensor:
        """Replica row of each id (owner-major)."""
        ids = ids.to(torch.int64)
        if self.table.mode == native.PART_HASH:
            return (ids % self.world) * self.rps + ids // self.world
        owner = torch.clamp(ids // self.table.div, max=self.world - 1)
        return owner * self.rps + (ids - owner * self.table.div)

    def rows(self, ids: torch.Tensor) -> torch.Tensor:
        """Current replica values of ``ids`` (tests / debugging); ids of the worker's own shard are read
        from the master, which is what the training kernels use for them."""
        idx = self.row_index(ids)
        out = self.cache[idx, : self.table.dim]
        if self.own is not None:
            mine = (idx // self.rps) == self.own
            if bool(mine.any()):
                out = out.clone()
                out[mine] = self.table.local[idx[mine] - self.own * self.rps, : self.table.dim]
        return out

    @property
    def pending(self) -> torch.Tensor:
        """int64 ``[world]`` device counters: messages buffered per destination (fed by the bucket
        histogram kernel of the training step, consumed by the policy kernel)."""
        return self.state[: self.world]

    def reserve_total(self) -> int:
        """CTA slots the training kernel must leave free for the exchange running next to it."""
        everything_in_place = self.skip_mask == (1 << self.world) - 1
        return 0 if everything_in_place else self.n_ctas

    # -- policy ---------------------------------------------------------------------------------
    def _arm(self, n_records: int) -> None:
        """First micro-batch: derive the count threshold from the batch size and stagger the
        destinations ((o - rank) % sync_every micro-batches of head start)."""
        per_dest = max(1, int(n_records) // self.world)
        if self.sliced:
            self.flush_count = max(1, per_dest // 2)      # fires for every destination after every step
            self._armed = True
            return
        if self.flush_count is None:
            self.flush_count = max(1, int((self.sync_every - 0.5) * per_dest))
        if self.stagger and self.sync_every > 1:
            head = [((o - self.rank) % self.sync_every) * per_dest for o in range(self.world)]
            self.state[: self.world] += torch.tensor(head, dtype=torch.int64, device=self.state.device)
        self._armed = True

    def after_step(self, n_records: int, fed: bool = False) -> None:
        """Call once per micro-batch, BEFORE launching the training kernel of that micro-batch (the
        exchange CTAs then get their slots first).  ``fed``: the per-destination counters were already
        incremented on the device (``native.bucket_by_item(..., pending=replica.pending)``)."""
        if not self._armed:
            self._arm(n_records)
        dev = self.table.cuda_device
        cur = torch.cuda.current_stream(dev)
        ev = torch.cuda.Event()
        ev.record(cur)                      # counters of this micro-batch are visible after this point
        self.stream.wait_event(ev)
        with torch.cuda.stream(self.stream):
            native.flush_policy(self.state, self.world, count_max=self.flush_count,
                                interval_ns=self.interval_ns, require_all=self.require_all,
                                add_uniform=0 if fed else max(1, int(n_records) // self.world))
            if self._timing:
                t0 = torch.cuda.Event(enable_timing=True); t0.record(self.stream)
            native.replica_exchange(self.table.table_c, self.cache, self.base, state=self.state,
                                    n_ctas=self.n_ctas, n_stages=self.stages, sequential=self.sequential,
                                    slices=self.sync_every if self.sliced else 1,
                                    slice_offset=self.rank if self.stagger else 0, skip_mask=self.skip_mask)
            done = torch.cuda.Event(enable_timing=self._timing)
            done.record(self.stream)
            if self._timing:
                self._timed.append((t0, done))
        self._pending_events.append(done)
        if len(self._pending_events) > self.max_outstanding:   # bounded staleness / run-ahead
            cur.wait_event(self._pending_events.pop(0))
        self.steps += 1
        self._since_flush += 1
        METRICS.inc("replica_policy_evals")

    def exchange(self, wide: bool = True) -> None:
        """Flush every destination now (force) on the side stream; does not block the current stream."""
        dev = self.table.cuda_device
        cur = torch.cuda.current_stream(dev)
        ev = torch.cuda.Event()
        ev.record(cur)
        self.stream.wait_event(ev)
        with torch.cuda.stream(self.stream):
            native.flush_policy(self.state, self.world, force=True)
            n = 2 * native.sm_count(self.table.device) if wide else self.n_ctas
            native.replica_exchange(self.table.table_c, self.cache, self.base, state=self.state,
                                    n_ctas=n, n_stages=self.stages, skip_mask=self.skip_mask)
            done = torch.cuda.Event()
            done.record(self.stream)
        self._pending_events.append(done)
        self._since_flush = 0
        METRICS.inc("replica_exchanges")

    def flush(self) -> None:
        """Push every pending local delta to the masters and make the current stream wait for it."""
        if self._since_flush > 0 or self._pending_events:
            self.exchange()
        cur = torch.cuda.current_stream(self.table.cuda_device)
        for ev in self._pending_events:
            cur.wait_event(ev)
        self._pending_events = []

    def refresh(self) -> None:
        """Quiesce: after this (collective) call the master holds every rank's deltas and every
        replica equals the master."""
        self._since_flush = max(self._since_flush, 1)
        self.flush()
        self.table.barrier()
        self._since_flush = 1
        self.flush()
        self.table.barrier()

    def timing_summary(self):
        """``FPS_EXCHANGE_TIMING=1``: device time of the per-step exchange kernels (ms)."""
        if not self._timed:
            return None
        torch.cuda.synchronize(self.table.device)
        ms = sorted(a.elapsed_time(b) for a, b in self._timed)
        return {"n": len(ms), "median_ms": ms[len(ms) // 2], "p90_ms": ms[int(len(ms) * 0.9)], "max_ms": ms[-1]}

    def flush_counts(self) -> List[int]:
        """Number of flushes per destination so far (device counters)."""
        lo = native.FLUSH_COUNT
        return self.state[lo: lo + self.world].tolist()
