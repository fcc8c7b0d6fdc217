"""Comparison B of BASELINE.md: the same pull / push done with collectives + separate kernels.

``NcclOnlineMF`` has the API of :class:`DeviceOnlineMF` but implements one micro-batch as

    bucket item ids by owner -> all_to_all(ids) -> owners gather rows -> all_to_all(values)
    -> elementwise SGD kernels -> all_to_all(deltas) -> owners index_add_

i.e. exactly "a path that only calls NCCL for pull/push" with stock PyTorch kernels in between:
the baseline the fused one-sided kernels are measured against (``bench.py --impl nccl``).  With
the gloo backend and CPU tensors the same class is the multi-process CPU plumbing path
(BASELINE.json config 1: async SGD MF k=16, world_size=2, no GPU).
"""
from __future__ import annotations

from typing import Iterable, Optional, Sequence

import torch
import torch.distributed as dist

ERR_SIGMOID, ERR_PLAIN = 0, 1


def _uniform_by_id(ids: torch.Tensor, dim: int, seed: int, lo: float, hi: float) -> torch.Tensor:
    """Deterministic per-id init for the baseline (hash-mixed LCG; not the Philox of the kernels)."""
    j = torch.arange(dim, device=ids.device, dtype=torch.int64)[None, :]
    x = (ids[:, None].to(torch.int64) * 0x9E3779B1 + j * 0x85EBCA77 + seed * 0xC2B2AE3D) & 0x7FFFFFFF
    x = (x * 1103515245 + 12345) & 0x7FFFFFFF
    x = (x ^ (x >> 13)) * 0x5BD1E995 & 0x7FFFFFFF
    u = (x >> 7).to(torch.float32) / float(1 << 24)
    return lo + (hi - lo) * u


class NcclOnlineMF:
    def __init__(self, num_users: int, num_items: int, num_factors: int = 10,
                 range_min: float = -0.01, range_max: float = 0.01, learning_rate: float = 0.01,
                 negative_sample_rate: int = 0, pull_limit: int = 0, group=None, seed: int = 0,
                 err_mode: int = ERR_SIGMOID, device: Optional[torch.device] = None, **_ignored):
        ready = dist.is_available() and dist.is_initialized()
        self.group = group
        self.world = dist.get_world_size(group) if ready else 1
        self.rank = dist.get_rank(group) if ready else 0
        if device is None:
            device = (torch.device("cuda", torch.cuda.current_device())
                      if torch.cuda.is_available() else torch.device("cpu"))
        self.dev = torch.device(device)
        self.k, self.lr, self.err_mode = int(num_factors), float(learning_rate), int(err_mode)
        self.num_users, self.num_items = int(num_users), int(num_items)
        W = self.world
        n_items_local = -(-self.num_items // W)
        n_users_local = -(-self.num_users // W)
        item_ids = torch.arange(n_items_local, device=self.dev) * W + self.rank
        user_ids = torch.arange(n_users_local, device=self.dev) * W + self.rank
        self.item_shard = _uniform_by_id(item_ids, self.k, seed * 2 + 1, range_min, range_max)
        self.users = _uniform_by_id(user_ids, self.k, seed * 2 + 2, range_min, range_max)
        self.stats = torch.zeros(2, dtype=torch.float32, device=self.dev)
        self.launches = 0

    # -- collectives ------------------------------------------------------------------------
    def _a2a(self, send: torch.Tensor, send_counts, recv_counts) -> torch.Tensor:
        if self.world == 1:
            return send
        out = send.new_empty((int(sum(recv_counts)),) + tuple(send.shape[1:]))
        dist.all_to_all_single(out, send, output_split_sizes=list(recv_counts),
                               input_split_sizes=list(send_counts), group=self.group)
        return out

    def step(self, users: torch.Tensor, items: torch.Tensor, ratings: torch.Tensor) -> None:
        W = self.world
        items64 = items.to(torch.int64)
        owner = items64 % W
        order = torch.argsort(owner, stable=True)
        send_ids = items64[order]
        send_counts = torch.bincount(owner, minlength=W)
        if W > 1:
            recv_counts = torch.empty_like(send_counts)
            dist.all_to_all_single(recv_counts, send_counts, group=self.group)
            sc, rc = send_counts.tolist(), recv_counts.tolist()
        else:
            sc = rc = send_counts.tolist()
        # PULL: request ids -> owners gather -> answers
        req = self._a2a(send_ids, sc, rc)
        rows = self.item_shard[req // W]
        v_sorted = self._a2a(rows, rc, sc)
        # worker compute (separate elementwise kernels)
        uslot = users.to(torch.int64)[order] // W
        u = self.users[uslot]
        r = ratings[order]
        resid = r - (u * v_sorted).sum(1)
        e = torch.sigmoid(resid) if self.err_mode == ERR_SIGMOID else resid
        g = (self.lr * e)[:, None]
        self.users.index_add_(0, uslot, g * v_sorted)
        dv = g * u
        # PUSH: deltas -> owners -> paramUpdate
        dv_recv = self._a2a(dv, sc, rc)
        self.item_shard.index_add_(0, req // W, dv_recv)
        self.stats[0] += (resid * resid).sum()
        self.stats[1] += float(resid.numel())

    def fit_stream(self, host_batches: Iterable[Sequence[torch.Tensor]]):
        for (u, i, r) in host_batches:
            self.stats.zero_()
            self.step(u.to(self.dev, non_blocking=True), i.to(self.dev, non_blocking=True),
                      r.to(self.dev, non_blocking=True))
            s = self.stats.cpu()
            yield float(s[0]), float(s[1])

    def predict(self, users: torch.Tensor, items: torch.Tensor) -> torch.Tensor:
        W = self.world
        items64 = items.to(torch.int64)
        owner = items64 % W
        order = torch.argsort(owner, stable=True)
        send_counts = torch.bincount(owner, minlength=W)
        if W > 1:
            recv_counts = torch.empty_like(send_counts)
            dist.all_to_all_single(recv_counts, send_counts, group=self.group)
            sc, rc = send_counts.tolist(), recv_counts.tolist()
        else:
            sc = rc = send_counts.tolist()
        req = self._a2a(items64[order], sc, rc)
        v = self._a2a(self.item_shard[req // W], rc, sc)
        u = self.users[users.to(torch.int64)[order] // W]
        out = torch.empty(items.numel(), dtype=torch.float32, device=self.dev)
        out[order] = (u * v).sum(1)
        return out

    def check_finite(self) -> None:
        if not torch.isfinite(self.item_shard).all():
            raise FloatingPointError("non-finite item factors")

    def barrier(self) -> None:
        if self.dev.type == "cuda":
            torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier(group=self.group)

    def close(self) -> None:
        pass
