"""Synthetic rating streams with learnable structure (there is no network for datasets).

``lowrank_ratings(users, items)`` evaluates a fixed rank-``rank`` ground-truth model whose factors are
pseudo-random *functions of the id* (no tables, so 10M x 1M problems cost nothing to describe):

    r(u, i) = scale / sqrt(rank) * sum_f a_f(u) * b_f(i),      a_f, b_f ~ U(-1, 1) hashed from (id, f, seed)

Used by the convergence gates (tests/mp_replica_check.py, bench.py ``config.quality``): every parallel
mode must reach the same held-out RMSE as the single-worker run on the same update budget.
"""
from __future__ import annotations

import torch

_M1, _M2 = 0x9E3779B97F4A7C15 - (1 << 64), 0xBF58476D1CE4E5B9 - (1 << 64)   # as signed int64


def _mix(x: torch.Tensor) -> torch.Tensor:
    """splitmix64-style finaliser on int64 tensors (wrap-around arithmetic)."""
    x = (x ^ ((x >> 30) & 0x3FFFFFFFF)) * _M2
    x = (x ^ ((x >> 27) & 0x1FFFFFFFFF)) * _M1
    return x ^ ((x >> 31) & 0x1FFFFFFFF)


def hashed_uniform(ids: torch.Tensor, f: int, seed: int) -> torch.Tensor:
    """U(-1, 1) as a pure function of (id, f, seed); fp32."""
    x = _mix(ids.to(torch.int64) * _M1 + (f + 1) * 0x632BE5AB + seed * 0x1B873593)
    return ((x >> 40) & 0xFFFFFF).to(torch.float32) * (2.0 / 16777216.0) - 1.0


def lowrank_ratings(users: torch.Tensor, items: torch.Tensor, rank: int = 8, seed: int = 0,
                    scale: float = 1.5) -> torch.Tensor:
    r = torch.zeros(users.shape, dtype=torch.float32, device=users.device)
    for f in range(rank):
        r += hashed_uniform(users, f, 2 * seed + 1) * hashed_uniform(items, f, 2 * seed + 2)
    return r * (scale / rank ** 0.5)
