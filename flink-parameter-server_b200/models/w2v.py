"""word2vec skip-gram with negative sampling on the parameter server (BASELINE.json config 4:
"word2vec skip-gram negative-sampling dim=300, push fused with paramUpdate").

Both embedding matrices live on the PS (``W_in`` and ``W_out``, sharded ``word % G``); a worker
step over a micro-batch of (center, context) pairs is ONE launch of the fused pull+SGD+push kernel
(ops/csrc/fps_core.cu with ``err_mode = 2``: ``g = lr * (label - sigmoid(u.v))``): it pulls both rows
from their owners, and pushes ``g*v`` to ``W_in[center]`` and ``g*u`` to ``W_out[context]`` with
``red.global.add.v4.f32`` -- the additive paramUpdate happens in the owner's memory system.
``negative`` extra pairs per positive are sampled on the device (uniform over the vocabulary) with
label 0.  Not part of the reference's algorithm suite; it exercises the same API on a 1.2 KB row.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from ..errors import FactorIsNotANumberException
from ..ops import native
from ..store.replica_cache import ReplicaCache
from ..store.sharded_table import ShardedTable

ERR_LOGISTIC = 2


class DeviceSkipGram:
    def __init__(self, vocab: int, dim: int = 300, learning_rate: float = 0.025, negative: int = 5,
                 group=None, seed: int = 0, device: Optional[int] = None,
                 replica_cache: Optional[bool] = None, sync_every: int = 4):
        self.vocab, self.dim, self.lr, self.negative, self.seed = vocab, dim, learning_rate, negative, seed
        b = 0.5 / dim
        self.w_in = ShardedTable(vocab, dim, group=group, device=device, init_range=(-b, b), seed=2 * seed + 1)
        self.w_out = ShardedTable(vocab, dim, group=group, device=device, init="zeros")
        self.dev = self.w_in.cuda_device
        self.stats = torch.zeros(2, dtype=torch.float32, device=self.dev)
        self.nan_flag = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.step_no = 0
        # multi-GPU: train local replicas of both tables and exchange deltas in the background
        # (store/replica_cache.py) instead of moving 2 x 1200 B per update over NVLink
        if replica_cache is None:
            replica_cache = self.w_in.world > 1
        self.rep_in = ReplicaCache(self.w_in, sync_every) if replica_cache else None
        self.rep_out = ReplicaCache(self.w_out, sync_every) if replica_cache else None
        self._ones = None

    def step(self, centers: torch.Tensor, contexts: torch.Tensor) -> None:
        if self._ones is None or self._ones.numel() != centers.numel():
            self._ones = torch.ones(centers.numel(), dtype=torch.float32, device=self.dev)
        tin = self.rep_in.table_c if self.rep_in else self.w_in.table_c
        tout = self.rep_out.table_c if self.rep_out else self.w_out.table_c
        if self.rep_in:   # policy + exchange kernels first (side streams), then the training kernel
            n = centers.numel() * (1 + self.negative)
            self.rep_in.after_step(n); self.rep_out.after_step(n)
        native.mf_sgd_fused(centers, contexts, self._ones, tin, 1, tout, self.lr,
                            err_mode=ERR_LOGISTIC, neg_rate=self.negative, num_items=self.vocab,
                            seed=self.seed, step=self.step_no, stats=self.stats, nan_flag=self.nan_flag,
                            kernel="reg",
                            reserve_total=(self.rep_in.reserve_total() + self.rep_out.reserve_total())
                            if self.rep_in else 0)
        self.step_no += 1

    def flush(self) -> None:
        if self.rep_in:
            self.rep_in.flush(); self.rep_out.flush()

    def similarity(self, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        self.flush()
        va, vb = self.w_in.pull(a), self.w_in.pull(b)
        return torch.nn.functional.cosine_similarity(va, vb)

    def score(self, centers: torch.Tensor, contexts: torch.Tensor) -> torch.Tensor:
        """sigmoid(W_in[center] . W_out[context]) -- the model's co-occurrence probability."""
        self.flush()
        u = self.w_in.pull(centers)
        return torch.sigmoid(self.w_out.pull_dot(contexts, torch.nn.functional.pad(
            u, (0, self.w_out.stride - u.shape[1])).contiguous()))

    def check_finite(self):
        if int(self.nan_flag.item()):
            raise FactorIsNotANumberException("non-finite skip-gram update")

    def barrier(self):
        self.flush()
        self.w_in.barrier()

    def close(self):
        self.w_in.close(); self.w_out.close()
