"""Comparison B of BASELINE.md: the same pull / push done with collectives + separate kernels.

``NcclOnlineMF`` has the API of :class:`DeviceOnlineMF` but implements one micro-batch as

    sort item ids by owner into fixed-capacity slots -> all_to_all(ids) -> owners gather rows ->
    all_to_all(values) -> elementwise SGD kernels -> all_to_all(deltas) -> owners index_add_
    (equal-split collectives: no size exchange, no host synchronisation inside a step)

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
        self.overflow = torch.zeros((), dtype=torch.bool, device=self.dev)
        self._n_cap = 0
        self.launches = 0

    # -- collectives ------------------------------------------------------------------------
    def _a2a(self, send: torch.Tensor, send_counts, recv_counts) -> torch.Tensor:
        if self.world == 1:
            return send
        out = send.new_empty((int(sum(recv_counts)),) + tuple(send.shape[1:]))
        dist.all_to_all_single(out, send, output_split_sizes=list(recv_counts),
                               input_split_sizes=list(send_counts), group=self.group)
        return out

    def _slots(self, n: int) -> int:
        """Fixed per-destination capacity of the exchange buffers: mean + 6 sigma of a uniform split, so the
        collectives need no size exchange (no host synchronisation per step); overflow is flagged."""
        W = self.world
        if W == 1:
            return n
        if n > self._n_cap:       # first step (or a larger batch): agree on the capacity once, collectively
            t = torch.tensor([n], dtype=torch.int64, device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            self._n_cap = int(t.item())
        mean = self._n_cap / W
        return int(mean + 6.0 * (mean * (1.0 - 1.0 / W)) ** 0.5 + 16)

    def step(self, users: torch.Tensor, items: torch.Tensor, ratings: torch.Tensor) -> None:
        """One micro-batch with collectives + stock kernels and NO host synchronisation:
        owner-sort -> fixed-capacity send slots -> all_to_all(ids) -> gather -> all_to_all(rows)
        -> elementwise SGD -> index_add (users) -> all_to_all(deltas) -> index_add (item shard)."""
        W, k = self.world, self.k
        n = items.numel()
        C = self._slots(n)
        items64 = items.to(torch.int64)
        owner = items64 % W
        order = torch.argsort(owner, stable=True)
        so = owner[order]
        counts = torch.bincount(owner, minlength=W)
        start = torch.cumsum(counts, 0) - counts
        pos = torch.arange(n, device=self.dev) - start[so]                 # rank inside the owner group
        self.overflow |= (counts.max() > C)
        keep = pos < C
        slot = (so * C + pos)[keep]
        src = order[keep]
        send_ids = torch.full((W * C,), -1, dtype=torch.int64, device=self.dev)
        send_ids[slot] = items64[src]
        # PULL: request ids -> owners gather -> answers
        req = self._a2a_eq(send_ids)
        valid_req = req >= 0
        rows = self.item_shard[torch.where(valid_req, req // W, torch.zeros_like(req))]
        v = self._a2a_eq(rows)                                            # [W*C, k] answers, slot order
        # worker compute (separate elementwise kernels) on the padded slots
        uslot = torch.zeros(W * C, dtype=torch.int64, device=self.dev)
        uslot[slot] = users.to(torch.int64)[src] // W
        r = torch.zeros(W * C, dtype=torch.float32, device=self.dev)
        r[slot] = ratings[src]
        live = send_ids >= 0
        u = self.users[uslot]
        resid = r - (u * v).sum(1)
        e = torch.sigmoid(resid) if self.err_mode == ERR_SIGMOID else resid
        g = torch.where(live, self.lr * e, torch.zeros_like(e))[:, None]
        self.users.index_add_(0, uslot, g * v)
        dv = g * u
        # PUSH: deltas -> owners -> paramUpdate
        dv_recv = self._a2a_eq(dv)
        self.item_shard.index_add_(0, torch.where(valid_req, req // W, torch.zeros_like(req)),
                                   torch.where(valid_req[:, None], dv_recv, torch.zeros_like(dv_recv)))
        live_f = live.float()
        self.stats[0] += (resid * resid * live_f).sum()
        self.stats[1] += live_f.sum()

    def _a2a_eq(self, send: torch.Tensor) -> torch.Tensor:
        """all_to_all with equal splits (no size exchange)."""
        if self.world == 1:
            return send
        out = torch.empty_like(send)
        dist.all_to_all_single(out, send, group=self.group)
        return out

    def fit_stream(self, host_batches: Iterable[Sequence[torch.Tensor]]):
        """Every batch copied from (pinned) host memory, the step's loss read back one step late."""
        pending = []
        for (u, i, r) in host_batches:
            self.stats.zero_()
            self.step(u.to(self.dev, non_blocking=True), i.to(self.dev, non_blocking=True),
                      r.to(self.dev, non_blocking=True))
            if self.dev.type == "cuda":
                h = torch.empty(2, dtype=torch.float32).pin_memory()
                h.copy_(self.stats, non_blocking=True)
                ev = torch.cuda.Event(); ev.record()
                pending.append((h, ev))
                if len(pending) > 2:
                    h0, e0 = pending.pop(0); e0.synchronize()
                    yield float(h0[0]), float(h0[1])
            else:
                s = self.stats.clone()
                yield float(s[0]), float(s[1])
        for h0, e0 in pending:
            e0.synchronize()
            yield float(h0[0]), float(h0[1])

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
        if bool(self.overflow):
            raise RuntimeError("NCCL baseline: an exchange slot overflowed (raise the slot capacity)")
        if not torch.isfinite(self.item_shard).all():
            raise FloatingPointError("non-finite item factors")

    def barrier(self) -> None:
        if self.dev.type == "cuda":
            torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier(group=self.group)

    def close(self) -> None:
        pass
