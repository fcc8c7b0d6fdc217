"""Device (B200) implementation of online / offline SGD matrix factorisation.

Capability parity with ``PSOnlineMatrixFactorization.psOnlineMF`` and
``PSOfflineMatrixFactorization.psOfflineMF`` (reference:
M/matrix/factorization/PSOnlineMatrixFactorization.scala:39-75,
M/matrix/factorization/workers/PSOnlineMatrixFactorizationWorker.scala:22-90):

* user vectors live on the *worker* that owns the user (``user % workerParallelism``),
* item vectors live on the parameter server, sharded ``item % psParallelism``,
* every rating triggers pull(item) -> SGD delta -> local user update -> push(item delta).

B200-first mechanism: one process per GPU is both worker ``rank`` and PS shard ``rank``; the whole
worker step for a micro-batch is ONE kernel (``fps_mf_sgd_fused``) that pulls item rows with
16-byte loads from the owner's HBM over NVSwitch, computes the update and pushes the delta back
with ``red.global.add.v4.f32`` -- no messages, no NCCL, no separate elementwise kernel.
"""
from __future__ import annotations

import os
from typing import Iterable, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from ...errors import FactorIsNotANumberException
from ...ops import native
from ...utils.metrics import GLOBAL as METRICS
from ...runtime.device_stream import DevicePrefetcher
from ...store.replica_cache import ReplicaCache
from ...store.sharded_table import ShardedTable

ERR_SIGMOID = 0  # reference parity: e = sigmoid(r - u.v)   (SGDUpdater.scala:8)
ERR_PLAIN = 1    # textbook SGD:     e = r - u.v

DEFAULT_DEVICE_PULL_LIMIT = 0  # 0 = as many row slots in flight as the GPU can hold


class DeviceOnlineMF:
    def __init__(self, num_users: int, num_items: int, num_factors: int = 10,
                 range_min: float = -0.01, range_max: float = 0.01, learning_rate: float = 0.01,
                 negative_sample_rate: int = 0, pull_limit: int = DEFAULT_DEVICE_PULL_LIMIT,
                 group=None, seed: int = 0, err_mode: int = ERR_SIGMOID,
                 device: Optional[int] = None, track_touched: bool = False,
                 kernel: Optional[str] = None, item_cache: Optional[bool] = None,
                 sync_every: int = 4, user_memory: int = 0,
                 sync_interval_ms: Optional[float] = None, item_blocking: Optional[bool] = None,
                 block_bytes: int = 16 << 20, flush_count: Optional[int] = None,
                 flush_require: str = "any", replica_own_inplace: Optional[bool] = None,
                 output_ring=None):
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.cuda_device = torch.device("cuda", self.device)
        self.group = group
        ready = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if ready else 1
        self.rank = dist.get_rank(group) if ready else 0
        self.num_users, self.num_items, self.k = int(num_users), int(num_items), int(num_factors)
        self.lr = float(learning_rate)
        self.neg = int(negative_sample_rate)
        self.pull_limit = int(pull_limit)
        self.err_mode = int(err_mode)
        self.seed = int(seed)
        self.step_no = 0
        self.kernel = kernel
        # pull limiter (WL:196-250) = device credit counter [credits, stalls] consumed inside the fused kernel
        # (a warp takes its credits all at once, so limits below one warp's worth of pulls fall back to the
        # static form of the limiter: a capped grid)
        self.credits = (torch.tensor([self.pull_limit, 0], dtype=torch.int32, device=self.cuda_device)
                        if self.pull_limit >= 32 and os.environ.get("FPS_STATIC_LIMITER", "0") != "1" else None)
        self.output_ring = output_ring     # E5: per-update (user, vector) output stream (runtime/output_ring.py)
        with torch.cuda.device(self.device):
            # parameter server: item vectors, sharded item % psParallelism
            self.items = ShardedTable(num_items, num_factors, partition="hash", group=group,
                                      device=self.device, init="uniform",
                                      init_range=(range_min, range_max), seed=seed * 2 + 1,
                                      track_touched=track_touched)
            # worker-local state: vectors of the users this worker owns (user % W == rank)
            n_local = -(-self.num_users // self.world)
            self.users = torch.empty((n_local, self.items.stride), dtype=torch.float32,
                                     device=self.cuda_device)
            native.init_rows(self.users, self.k, self.rank, self.world, native.PART_HASH, n_local,
                             seed * 2 + 2, range_min, range_max)
            self.stats = torch.zeros(2, dtype=torch.float32, device=self.cuda_device)
            self.nan_flag = torch.zeros(1, dtype=torch.int32, device=self.cuda_device)
            # K5: per-user memory of recently seen items (userMemory of the reference, default 128
            # there; 0 here = sample uniformly inside the fused kernel, rejecting only the positive)
            self.user_memory = int(user_memory) if self.neg > 0 else 0
            if self.user_memory > 0:
                self.seen = torch.full((n_local, self.user_memory), -1, dtype=torch.int32,
                                       device=self.cuda_device)
                self.seen_pos = torch.zeros(n_local, dtype=torch.int32, device=self.cuda_device)
        # ---- item-cache mode (sender-side combining) --------------------------------------------
        # The worker trains a local owner-major replica of the item table (pulls and pushes stay in local
        # HBM); its segments are the per-destination send buffers of the reference's batching senders.
        # After every micro-batch a device-side count / timer policy (fps_flush_policy) picks the
        # destinations to flush -- by default each destination once per `sync_every` micro-batches,
        # staggered -- and a few TMA-driven CTAs (fps_replica_exchange) push (replica - base) to those
        # master shards and fold the other workers' contributions (master - base) into the replica while
        # the training kernel keeps running.  Asynchronous (no barriers); staleness ~ `sync_every` steps.
        if item_cache is None:
            item_cache = self.world > 1 and os.environ.get("FPS_ITEM_CACHE", "1") != "0"
        self.item_cache = bool(item_cache)
        self.sync_every = max(1, int(sync_every))
        self.sync_interval_ms = sync_interval_ms
        self.flush_count, self.flush_require = flush_count, flush_require
        self._own_inplace = replica_own_inplace
        self.replica = (ReplicaCache(self.items, self.sync_every, sync_interval_ms,
                                     require=flush_require, flush_count=flush_count,
                                     own_inplace=replica_own_inplace)
                        if self.item_cache else None)
        # ---- L2 blocking: deal each micro-batch into buckets of <= 16 MB of item rows (fps_bucket.cu) ----
        # only where the item rows are read from local HBM (single GPU, or the local replica)
        row_bytes = self.items.stride * 4
        if item_blocking is None:
            item_blocking = ((self.world == 1 or self.item_cache) and self.num_items * row_bytes > (48 << 20)
                             and (self.neg == 0 or self.user_memory > 0)
                             and os.environ.get("FPS_ITEM_BLOCKING", "1") != "0")
        self.item_blocking = bool(item_blocking)
        block_bytes = int(os.environ.get("FPS_BLOCK_BYTES", block_bytes))
        self.l2_hints = self.item_blocking and os.environ.get("FPS_L2_HINTS", "0") == "1"
        per_bucket = max(1, int(block_bytes) // row_bytes)
        self.block_shift = max(0, per_bucket.bit_length() - 1)
        # buckets are ranges of rows of the table the fused kernel reads: the item shard (N = 1) or the
        # owner-major replica (row = owner * rows_per_shard + slot)
        table_rows = self.items.rows_per_shard * self.world if self.item_cache else self.num_items
        while -(-table_rows >> self.block_shift) > native.BUCKET_MAX:
            self.block_shift += 1
        self.block_buckets = max(1, -(-table_rows >> self.block_shift))
        if self.item_blocking:
            with torch.cuda.device(self.device):
                self._bucket_scratch = torch.zeros(2 * native.BUCKET_MAX, dtype=torch.int32,
                                                   device=self.cuda_device)
        self.items.barrier()

    # ------------------------------------------------------------------------------------
    def flush(self) -> None:
        """Item-cache mode: push every pending local delta to the master shards and wait for it."""
        if self.replica is not None:
            self.replica.flush()

    def step(self, users: torch.Tensor, items: Optional[torch.Tensor] = None,
             ratings: Optional[torch.Tensor] = None) -> None:
        """Process one micro-batch of ratings whose users belong to this worker (async SGD).
        ``step(packed)`` with a single int64 tensor takes packed64 records (``native.pack_ratings``)."""
        neg = self.neg
        if self.user_memory > 0:
            # negatives drawn by the sampler kernel against the per-user seen ring; the fused kernel
            # then consumes the expanded batch as plain records
            users, items, ratings = native.neg_sample(users, items, ratings, self.neg, self.num_items,
                                                      self.seen, self.seen_pos, self.world,
                                                      seed=self.seed, step=self.step_no)
            neg = 0
        n_records = users.numel()
        fed = False
        ring = self.output_ring
        out_args = ring.kernel_args() if ring is not None else None
        if self.item_blocking and self.block_buckets > 1:
            hashed = self.item_cache and self.items.mode == native.PART_HASH
            fed = hashed
            users, items, ratings = native.bucket_by_item(
                users, items, ratings, self.block_shift, self.block_buckets, self._bucket_scratch,
                num_shards=self.world if hashed else 1, rows_per_shard=self.items.rows_per_shard,
                pending=self.replica.pending if hashed else None)
        if self.item_cache:
            # policy + exchange kernels of this micro-batch go first (side stream): their CTAs take the
            # slots the training grid leaves free
            self.replica.after_step(n_records * (1 + neg), fed=fed)
            native.mf_sgd_fused(users, items, ratings, self.users, self.world, self.replica.table_c,
                                self.lr, err_mode=self.err_mode, neg_rate=neg,
                                num_items=self.num_items, seed=self.seed, step=self.step_no,
                                stats=self.stats, nan_flag=self.nan_flag,
                                max_inflight_rows=self.pull_limit, kernel="reg", l2_hints=self.l2_hints,
                                reserve_total=self.replica.reserve_total(), output=out_args,
                                credits=self.credits)
        else:
            native.mf_sgd_fused(users, items, ratings, self.users, self.world, self.items.table_c,
                                self.lr, err_mode=self.err_mode, neg_rate=neg,
                                num_items=self.num_items, seed=self.seed, step=self.step_no,
                                stats=self.stats, nan_flag=self.nan_flag,
                                max_inflight_rows=self.pull_limit, kernel=self.kernel,
                                l2_hints=self.l2_hints, output=out_args, credits=self.credits)
        if ring is not None:               # device-side count / timer policy + flush to the pinned host ring
            ring.after_kernel(users.numel() * (1 + neg))
        self.step_no += 1
        METRICS.inc("mf_ratings", users.numel())

    def make_graph_step(self, batch_size: int, packed: bool = True):
        """CUDA-graph a fixed-size micro-batch step for launch-bound streaming (small batches).

        Returns ``(static_inputs, replay)``: copy the next batch into ``static_inputs`` (device
        tensors) and call ``replay()``; the captured graph contains the stats reset and the fused
        kernel, so one ``cudaGraphLaunch`` replaces the Python + ctypes launch path."""
        dev = self.cuda_device
        if packed:
            static = (torch.zeros(batch_size, dtype=torch.int64, device=dev),)
        else:
            static = (torch.zeros(batch_size, dtype=torch.int32, device=dev),
                      torch.zeros(batch_size, dtype=torch.int32, device=dev),
                      torch.zeros(batch_size, dtype=torch.float32, device=dev))
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(2):  # warm up outside capture
                self.stats.zero_(); self.step(*static)
        torch.cuda.current_stream(dev).wait_stream(s)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            self.stats.zero_()
            self.step(*static)
        return static, graph.replay

    def fit_stream(self, host_batches: Iterable[Sequence[torch.Tensor]],
                   loss_every: int = 1):
        """End-to-end training over pinned host micro-batches ``(users, items, ratings)``.

        Yields one host-side ``(sum_sq_err, n_updates)`` per micro-batch (device -> host read of
        the step's result), lagging the launch by one step so copies, kernels and reads overlap.
        """
        pf = DevicePrefetcher(host_batches, self.cuda_device, depth=2)
        self.prefetcher = pf
        pending = []
        ring = [torch.empty(2, dtype=torch.float32).pin_memory() for _ in range(4)]
        i = 0
        for batch in pf:
            self.stats.zero_()
            self.step(*batch)
            host = ring[i % len(ring)]
            host.copy_(self.stats, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            pending.append((host, ev))
            i += 1
            if len(pending) > 2:
                h, e = pending.pop(0)
                e.synchronize()
                yield float(h[0]), float(h[1])
        self.flush()
        for h, e in pending:
            e.synchronize()
            yield float(h[0]), float(h[1])

    # -- quality / export -------------------------------------------------------------------
    def predict(self, users: torch.Tensor, items: torch.Tensor) -> torch.Tensor:
        """u.v for (user, item) pairs whose users are local (pull fused with the dot)."""
        self.flush()
        slots = (users.to(torch.int64) // self.world)
        local = self.users[slots].contiguous()
        return self.items.pull_dot(items, local)

    def user_vectors(self) -> Tuple[torch.Tensor, torch.Tensor]:
        n_local = self.users.shape[0]
        ids = torch.arange(n_local, device=self.cuda_device) * self.world + self.rank
        sel = ids < self.num_users
        return ids[sel], self.users[sel, : self.k].clone()

    def item_vectors(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """All item vectors of the local shard (the fused kernel does not maintain the touched bitmap:
        with Philox lazy-init every id has a well-defined value whether or not it was pulled)."""
        self.flush()
        return self.items.dump_local(only_touched=False)

    # -- checkpoint / resume (the reference only has export + transformWithModelLoad) ------
    def save(self, directory: str) -> str:
        """Every rank writes its user partition and its item shard to ``directory/rank<r>.npz``."""
        import numpy as np

        os.makedirs(directory, exist_ok=True)
        self.barrier()
        uid, uvec = self.user_vectors()
        iid, ivec = self.item_vectors()
        path = os.path.join(directory, f"rank{self.rank}_of{self.world}.npz")
        np.savez(path, user_ids=uid.cpu().numpy(), user_vecs=uvec.cpu().numpy(), item_ids=iid.cpu().numpy(),
                 item_vecs=ivec.cpu().numpy(), step_no=self.step_no)
        return path

    def load(self, directory: str) -> None:
        """Resume from :meth:`save` (same world size): users go back to their worker, items to their
        shard (one-sided assign), replicas are re-pulled."""
        import numpy as np

        d = np.load(os.path.join(directory, f"rank{self.rank}_of{self.world}.npz"))
        uid = torch.from_numpy(d["user_ids"]).to(self.cuda_device)
        self.users[uid // self.world, : self.k] = torch.from_numpy(d["user_vecs"]).to(self.cuda_device)
        self.items.load(torch.from_numpy(d["item_ids"]).to(self.cuda_device),
                        torch.from_numpy(d["item_vecs"]).to(self.cuda_device))
        self.step_no = int(d["step_no"])
        self.items.barrier()
        if self.replica is not None:
            self.replica = ReplicaCache(self.items, self.sync_every, self.sync_interval_ms,
                                        require=self.flush_require, flush_count=self.flush_count,
                                        own_inplace=self._own_inplace)

    def check_finite(self) -> None:
        if int(self.nan_flag.item()) != 0:
            raise FactorIsNotANumberException("non-finite SGD update")

    def barrier(self) -> None:
        self.flush()
        self.items.barrier()

    def refresh(self) -> None:
        """Collective quiesce: every delta is in the masters and every replica equals the master."""
        if self.replica is not None:
            self.replica.refresh()
        else:
            self.items.barrier()

    def close(self) -> None:
        self.items.close()
