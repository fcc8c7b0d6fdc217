"""fp64 online SGD matrix factorisation on the device tier -- the like-for-like precision of the
reference, whose factor vectors are ``Array[Double]`` (M/matrix/factorization/utils/Vector.scala:8).

Same placement and mechanism as :class:`~fps_b200.models.mf.device.DeviceOnlineMF` in its *direct* mode
(user rows on the owning worker, item rows on the PS shards, one fused pull + SGD + push kernel per
micro-batch, ``ops/csrc/fps_mf_f64.cu``: ``ld.global.v2.f64`` pulls, ``red.global.add.f64`` pushes).  Rows
are twice as wide as in fp32, so the bandwidth-bound step runs at about half the updates/s.  L2 blocking
of the micro-batch is applied on a single GPU; the replica mode, negative sampling and the output stream
are fp32-tier features.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

from ...errors import FactorIsNotANumberException
from ...ops import native
from ...store.sharded_table import ShardedTable

ERR_SIGMOID, ERR_PLAIN = 0, 1


class DeviceOnlineMFf64:
    def __init__(self, num_users: int, num_items: int, num_factors: int = 10, range_min: float = -0.01,
                 range_max: float = 0.01, learning_rate: float = 0.01, group=None, seed: int = 0,
                 err_mode: int = ERR_SIGMOID, device: Optional[int] = None, block_bytes: int = 16 << 20):
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.cuda_device = torch.device("cuda", self.device)
        ready = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if ready else 1
        self.rank = dist.get_rank(group) if ready else 0
        self.num_users, self.num_items, self.k = int(num_users), int(num_items), int(num_factors)
        self.lr, self.err_mode = float(learning_rate), int(err_mode)
        with torch.cuda.device(self.device):
            # one row = k doubles = 2k four-byte cells of a ShardedTable
            self.items = ShardedTable(num_items, 2 * self.k, group=group, device=self.device, init="zeros")
            self.kd = self.items.stride // 2                       # doubles per (padded) row
            self.items_f64 = self.items.local.view(torch.float64)  # [rows_per_shard, kd]
            native.init_rows_f64(self.items_f64, self.k, self.rank, self.world, native.PART_HASH,
                                 self.items.rows_per_shard, seed * 2 + 1, range_min, range_max)
            n_local = -(-self.num_users // self.world)
            self.users = torch.empty((n_local, self.kd), dtype=torch.float64, device=self.cuda_device)
            native.init_rows_f64(self.users, self.k, self.rank, self.world, native.PART_HASH, n_local,
                                 seed * 2 + 2, range_min, range_max)
            self.stats = torch.zeros(2, dtype=torch.float32, device=self.cuda_device)
            self.nan_flag = torch.zeros(1, dtype=torch.int32, device=self.cuda_device)
        row_bytes = self.items.stride * 4
        self.item_blocking = self.world == 1 and self.num_items * row_bytes > (48 << 20)
        per_bucket = max(1, int(block_bytes) // row_bytes)
        self.block_shift = max(0, per_bucket.bit_length() - 1)
        while -(-self.num_items >> self.block_shift) > native.BUCKET_MAX:
            self.block_shift += 1
        self.block_buckets = max(1, -(-self.num_items >> self.block_shift))
        if self.item_blocking:
            self._scratch = torch.zeros(2 * native.BUCKET_MAX, dtype=torch.int32, device=self.cuda_device)
        self.items.barrier()

    def step(self, users: torch.Tensor, items: Optional[torch.Tensor] = None,
             ratings: Optional[torch.Tensor] = None) -> None:
        """One micro-batch of ratings whose users belong to this worker; ``step(packed)`` takes packed64."""
        if self.item_blocking and self.block_buckets > 1:
            users, items, ratings = native.bucket_by_item(users, items, ratings, self.block_shift,
                                                          self.block_buckets, self._scratch)
        native.mf_sgd_fused_f64(users, items, ratings, self.users, self.world, self.items.table_c, self.lr,
                                err_mode=self.err_mode, stats=self.stats, nan_flag=self.nan_flag)

    def predict(self, users: torch.Tensor, items: torch.Tensor) -> torch.Tensor:
        raw = self.items.pull(items)                                   # [n, 2k] four-byte cells
        v = raw.contiguous().view(torch.float64)[:, : self.k]
        u = self.users[users.to(torch.int64) // self.world, : self.k]
        return (u * v).sum(1)

    def user_vectors(self) -> Tuple[torch.Tensor, torch.Tensor]:
        ids = torch.arange(self.users.shape[0], device=self.cuda_device) * self.world + self.rank
        sel = ids < self.num_users
        return ids[sel], self.users[sel, : self.k].clone()

    def item_vectors(self) -> Tuple[torch.Tensor, torch.Tensor]:
        self.items.barrier()
        ids = self.items.local_ids()
        sel = ids < self.num_items
        return ids[sel], self.items_f64[sel, : self.k].clone()

    def check_finite(self) -> None:
        if int(self.nan_flag.item()) != 0:
            raise FactorIsNotANumberException("non-finite SGD update")

    def barrier(self) -> None:
        self.items.barrier()

    def close(self) -> None:
        self.items.close()
