"""Device top-K recommendation (K6 / K12): pull query vectors from the PS, score them against the
worker-local item table on the tcgen05 tensor cores, keep an exact top-K.

Capability of ``psTopKGenerator`` / the generator half of ``psOnlineLearnerAndGenerator``
(PSTopKGenerator.scala:47-107, PSTopKGeneratorWorker.scala:35-114): user vectors live on the PS,
item vectors on the workers, every query is answered by every worker with its local top-``workerK``
and the partial lists are merged (``merge_partial_topk`` = CollectTopKFromEachWorker).

Algorithm (exact, tile-pruned -- the GPU-idiomatic replacement of the LEMP bucket scan):
  pass 1  tensor-core GEMM, epilogue keeps only the per-(query, 128-item tile) maximum;
  theta   K-th largest tile maximum per query  (a lower bound of the true K-th best score);
  pass 2  same GEMM, epilogue appends every (score, item) >= theta  (a few x K candidates);
  select  top-K of the candidates.
Scores are TF32 products accumulated in FP32; ``rescore=True`` recomputes the K winners in full
FP32 (ordering among near-ties may then differ from the TF32 ranking by < 1e-3 relative).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from ...ops import native
from ...store.sharded_table import ShardedTable


class DeviceTopK:
    """``sort_by_length=True`` adds the LEMP LENGTH bound at tile granularity (K6;
    LEMPPruningFunctions.scala:20-27, the bucket early-exit of PSTopKGeneratorWorker.scala:86-90): the
    item table is kept sorted by vector length (descending), so every 128-item tile has a known
    maximum length and a tile can contain a top-K item of query ``q`` only if
    ``maxLen(tile) * |q| >= theta_q``.  The tiles that survive form a *prefix* of the table, so pruning
    is simply running the same tensor-core kernel over fewer rows -- no pointer chasing."""

    LENGTH_SLACK = 1.004  # TF32 products may exceed the fp32 Cauchy-Schwarz bound by ~2^-10 relative

    def __init__(self, items: torch.Tensor, max_batch_bytes: int = 512 << 20, sort_by_length: bool = False):
        if items.dim() != 2 or items.shape[1] % 4 != 0:
            raise ValueError("items must be [n_items, stride] with stride % 4 == 0")
        self.perm = None
        if sort_by_length:
            lens = items.norm(dim=1)
            self.perm = torch.argsort(lens, descending=True)
            items = items[self.perm].contiguous()
            self.tile_maxlen = lens[self.perm][:: native.TOPK_TILE].contiguous()
        self.items = items
        self.n_items, self.stride = items.shape
        self.n_tiles = (self.n_items + native.TOPK_TILE - 1) // native.TOPK_TILE
        self.max_batch_bytes = max_batch_bytes
        self.last_tiles_scored = (0, 0)   # (pass 1, pass 2) tiles actually scored by the last topk()

    def _tiles_needed(self, theta: torch.Tensor, qnorm: torch.Tensor) -> int:
        """Number of leading tiles that can still hold a score >= theta for at least one query."""
        bound = torch.where((theta > 0) & (qnorm > 0), theta / qnorm.clamp_min(1e-30),
                            torch.full_like(theta, -1.0))          # theta <= 0: the bound cannot prune
        need = (self.tile_maxlen[None, :] * self.LENGTH_SLACK >= bound[:, None]).sum(1)
        return int(need.max().item())

    # -- raw scores (validation / tiny problems) ---------------------------------------------
    def scores(self, *, q_ids=None, q_table: Optional[ShardedTable] = None, q_local=None) -> torch.Tensor:
        n_q = q_ids.numel() if q_ids is not None else q_local.shape[0]
        out = torch.empty((n_q, self.n_items), dtype=torch.float32, device=self.items.device)
        native.topk_mma(self.items, 0, q_ids=q_ids, q_tab=q_table.table_c if q_table else None,
                        q_local=q_local, out_scores=out)
        return out

    def topk(self, K: int, *, q_ids=None, q_table: Optional[ShardedTable] = None, q_local=None,
             rescore: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
        """Returns ``(scores [n_q, K'], item_rows [n_q, K'])`` best first, ``K' = min(K, n_items)``;
        item_rows index the local item table."""
        n_q = q_ids.numel() if q_ids is not None else q_local.shape[0]
        Kp = min(K, self.n_items)
        # worst case K*128 candidates per query (K tiles reach theta); typically a few x K.  Start
        # with a small buffer and redo pass 2 with the worst-case capacity only if a row overflowed.
        cap_full = min(self.n_items, max(Kp * native.TOPK_TILE, native.TOPK_TILE))
        cap_small = min(cap_full, max(1024, 8 * Kp))
        chunk = max(128, (self.max_batch_bytes // (cap_full * 8)) // 128 * 128)
        outs, outi = [], []
        dev = self.items.device
        for a in range(0, n_q, chunk):
            b = min(n_q, a + chunk)
            ids = q_ids[a:b].contiguous() if q_ids is not None else None
            ql = q_local[a:b].contiguous() if q_local is not None else None
            tab = q_table.table_c if q_table is not None else None
            n = b - a
            T = native.TOPK_TILE
            if self.perm is None:
                p2 = self.n_tiles
                tile_max = torch.empty((n, self.n_tiles), dtype=torch.float32, device=dev)
                native.topk_mma(self.items, 1, q_ids=ids, q_tab=tab, q_local=ql, tile_max=tile_max)
                p1 = self.n_tiles
            else:
                # LENGTH-pruned pass 1: score the tiles of the longest items first; their K-th best tile
                # maximum already bounds how far down the length-sorted table a top-K item can sit.
                q = q_table.pull(ids) if ids is not None else ql[:, : self.stride]
                qnorm = q.norm(dim=1)
                p1 = min(self.n_tiles, max(Kp, 8, self.n_tiles // 8))
                tile_max = torch.empty((n, p1), dtype=torch.float32, device=dev)
                native.topk_mma(self.items[: p1 * T], 1, q_ids=ids, q_tab=tab, q_local=ql, tile_max=tile_max)
                if p1 >= Kp and p1 < self.n_tiles:
                    theta0 = torch.topk(tile_max, Kp, dim=1).values[:, -1].contiguous()
                    p_ext = self._tiles_needed(theta0, qnorm)
                else:
                    p_ext = self.n_tiles
                if p_ext > p1:
                    more = torch.empty((n, p_ext - p1), dtype=torch.float32, device=dev)
                    native.topk_mma(self.items[p1 * T: p_ext * T], 1, q_ids=ids, q_tab=tab, q_local=ql,
                                    tile_max=more)
                    tile_max = torch.cat([tile_max, more], 1)
                    p1 = p_ext
            if tile_max.shape[1] >= Kp:
                theta = torch.topk(tile_max, Kp, dim=1).values[:, -1].contiguous()
            else:
                theta = torch.full((n,), -3.0e38, dtype=torch.float32, device=dev)
            if self.perm is not None:
                p2 = min(p1, self._tiles_needed(theta, qnorm)) if tile_max.shape[1] >= Kp else self.n_tiles
            scored = self.items if p2 == self.n_tiles else self.items[: p2 * T]
            self.last_tiles_scored = (p1, p2)
            for cap in (cap_small, cap_full):
                cnt = torch.zeros(n, dtype=torch.int32, device=dev)
                cs = torch.empty((n, cap), dtype=torch.float32, device=dev)
                ci = torch.empty((n, cap), dtype=torch.int32, device=dev)
                native.topk_mma(scored, 2, q_ids=ids, q_tab=tab, q_local=ql, theta=theta,
                                cand_count=cnt, cand_score=cs, cand_item=ci)
                if cap == cap_full or int(cnt.max().item()) <= cap:
                    break
            valid = torch.arange(cap, device=dev)[None, :] < cnt.clamp(max=cap)[:, None]
            cs = torch.where(valid, cs, torch.full_like(cs, -3.0e38))
            top = torch.topk(cs, Kp, dim=1)
            rows = torch.gather(ci, 1, top.indices).to(torch.int64)
            sc = top.values
            if rescore:
                q = (q_table.pull(ids) if ids is not None else ql[:, : self.stride])
                q = torch.nn.functional.pad(q, (0, self.stride - q.shape[1]))
                exact = torch.einsum("qd,qkd->qk", q, self.items[rows])
                order = torch.argsort(exact, dim=1, descending=True)
                sc, rows = torch.gather(exact, 1, order), torch.gather(rows, 1, order)
            if self.perm is not None:
                rows = self.perm[rows]      # back to row numbers of the caller's (unsorted) table
            outs.append(sc); outi.append(rows)
        return torch.cat(outs), torch.cat(outi)


def merge_partial_topk(scores: torch.Tensor, items: torch.Tensor, K: int,
                       seen_mask: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """K-way merge of per-worker partial lists ``[n_q, W*workerK]`` with an optional seen-item filter
    (CollectTopKFromEachWorker.scala:41-56)."""
    if seen_mask is not None:
        scores = torch.where(seen_mask, torch.full_like(scores, -3.0e38), scores)
    top = torch.topk(scores, min(K, scores.shape[1]), dim=1)
    return top.values, torch.gather(items, 1, top.indices)


class DistributedTopK:
    """Top-K serving across ranks (capability of ``psTopKGenerator``): user vectors on the PS
    (``user_table``), every rank holds a partition of the items (``local_items`` with their global ids
    ``local_item_ids``); every query is answered by every rank with its local top-``workerK`` (the
    reference broadcasts each rating to all workers, PSTopKGenerator.scala:78-89) and the partial lists
    are merged (E9: gather + K-way merge = ``CollectTopKFromEachWorker``)."""

    def __init__(self, user_table: ShardedTable, local_items: torch.Tensor, local_item_ids: torch.Tensor,
                 group=None):
        import torch.distributed as dist

        self.users, self.group = user_table, group
        self.local = DeviceTopK(local_items)
        self.item_ids = local_item_ids.to(torch.int64)
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1

    def topk(self, query_user_ids: torch.Tensor, K: int, workerK: Optional[int] = None):
        import torch.distributed as dist

        wk = min(workerK or K, self.local.n_items)
        sc, rows = self.local.topk(wk, q_ids=query_user_ids, q_table=self.users)
        gids = self.item_ids[rows]
        if self.world == 1:
            return merge_partial_topk(sc, gids, K)
        if wk < (workerK or K):  # pad so all ranks contribute equally sized lists
            pad = (workerK or K) - wk
            sc = torch.nn.functional.pad(sc, (0, pad), value=-3.0e38)
            gids = torch.nn.functional.pad(gids, (0, pad), value=-1)
        all_sc = [torch.empty_like(sc) for _ in range(self.world)]
        all_id = [torch.empty_like(gids) for _ in range(self.world)]
        dist.all_gather(all_sc, sc.contiguous(), group=self.group)
        dist.all_gather(all_id, gids.contiguous(), group=self.group)
        return merge_partial_topk(torch.cat(all_sc, 1), torch.cat(all_id, 1), K)
