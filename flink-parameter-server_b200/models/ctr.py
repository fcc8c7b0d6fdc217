"""Wide-&-deep style CTR model with the embedding table on the parameter server (BASELINE.json
config 3: "wide-&-deep CTR (1B-slot embedding shard) async SGD, pull-limiter=64").

The sparse side (hashed categorical features -> embedding rows) is the parameter server: rows are
pulled with the one-sided gather kernel under the **device credit-counter pull limiter** -- at most
``pull_limit`` row pulls un-answered at any time, whatever the grid (``fps_pull_gather`` with credits,
WL:196-250) -- the dense tower (208-256-1 MLP + the "wide" linear term, replicated per worker) runs as ONE
hand-written fused kernel (forward, BCE, backward, dense SGD, row gradients; ``ops/csrc/fps_ctr.cu``: no
cuBLAS, no autograd, no host synchronisation), and the embedding gradients are pushed back with
``red.add`` (asynchronous SGD: no barrier between workers).  A step is 4 launches: gather, tower, apply,
push.  Not part of the reference's algorithm suite; it exercises the generic tensor tier at
embedding-table scale (1 B slots x 12 floats x 4 B = 48 GB = 6 GB per GPU at N = 8).
"""
from __future__ import annotations

from typing import Optional

import torch

from ..ops import native
from ..store.sharded_table import ShardedTable


class DeviceWideAndDeep:
    def __init__(self, num_slots: int, fields: int, emb_dim: int = 8, hidden: int = 256,
                 learning_rate: float = 0.05, pull_limit: int = 64, group=None, seed: int = 0,
                 device: Optional[int] = None):
        if hidden != native.CTR_HIDDEN:
            raise ValueError(f"the fused tower kernel is built for hidden = {native.CTR_HIDDEN}")
        if fields * emb_dim > 256:
            raise ValueError("fields * emb_dim must be <= 256")
        self.fields, self.emb_dim, self.lr, self.pull_limit = fields, emb_dim, learning_rate, pull_limit
        self.table = ShardedTable(num_slots, emb_dim + 1, group=group, device=device,
                                  init_range=(-0.05, 0.05), seed=seed)   # [embedding | wide weight]
        self.dev = self.table.cuda_device
        g = torch.Generator(device="cpu").manual_seed(seed)
        IN = fields * emb_dim
        W1 = (torch.randn(IN, hidden, generator=g) * 0.1).to(self.dev)
        self.w = {"W1": W1.contiguous(), "W1T": W1.t().contiguous(),
                  "b1": (torch.randn(hidden, generator=g) * 0.1).to(self.dev),
                  "w2": (torch.randn(hidden, generator=g) * 0.1).to(self.dev),
                  "b2": (torch.randn(1, generator=g) * 0.1).to(self.dev)}
        self.g = {k: torch.zeros_like(v) for k, v in self.w.items() if k != "W1T"}
        self.loss = torch.zeros(2, dtype=torch.float32, device=self.dev)   # [sum BCE, examples] of the last step
        self._rows = self._drows = None

    def _pull(self, ids: torch.Tensor) -> torch.Tensor:
        """One launch; the device credit counter keeps at most ``pull_limit`` row pulls in flight."""
        flat = ids.reshape(-1).contiguous()
        if self._rows is None or self._rows.shape[0] != flat.numel():
            self._rows = torch.empty((flat.numel(), self.table.stride), dtype=torch.float32, device=self.dev)
            self._drows = torch.empty_like(self._rows)
        native.pull_gather(self.table.table_c, flat, self._rows, max_inflight_rows=self.pull_limit,
                           credits=self.table._credits(self.pull_limit, self.dev) if self.pull_limit > 0 else None)
        return self._rows

    def step(self, ids: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        """ids: [B, fields] hashed feature slots, labels: [B] in {0,1}.  Returns the device tensor
        ``[sum of log-losses, examples]`` of this batch (no host synchronisation)."""
        rows = self._pull(ids)
        self.loss.zero_()
        native.ctr_step(rows, labels.float().contiguous(), self.fields, self.emb_dim, self.w, self.g,
                        self._drows, self.loss, self.lr, train=True)
        self.table.push(ids.reshape(-1), self._drows)          # d_rows is already -lr * gradient
        return self.loss

    def predict(self, ids: torch.Tensor) -> torch.Tensor:
        rows = self._pull(ids)
        prob = torch.empty(ids.shape[0], dtype=torch.float32, device=self.dev)
        scratch = torch.zeros(2, dtype=torch.float32, device=self.dev)
        native.ctr_step(rows, torch.zeros(ids.shape[0], device=self.dev), self.fields, self.emb_dim, self.w,
                        self.g, None, scratch, self.lr, train=False, prob=prob)
        return prob

    def credit_stalls(self) -> int:
        return self.table.credit_stalls()

    def close(self):
        self.table.close()
