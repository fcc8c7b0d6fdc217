"""Wide-&-deep style CTR model with the embedding table on the parameter server (BASELINE.json
config 3: "wide-&-deep CTR (1B-slot embedding shard) async SGD, pull-limiter=64").

The sparse side (hashed categorical features -> embedding rows) is the parameter server: rows are
pulled with the one-sided gather kernel under the device-side pull limiter (at most ``pull_limit``
row pulls in flight), the dense tower (a small MLP, replicated per worker) runs with torch autograd, and the
embedding gradients are pushed back with ``red.add`` (asynchronous SGD: no barrier between workers).
The "wide" linear term is an extra 1-wide column of the same rows.  Not part of the reference's
algorithm suite; it exercises the generic tensor tier at embedding-table scale
(slots x dim x 4 B per shard; 1 B slots x 8 floats = 32 GB of a 180 GB B200).
"""
from __future__ import annotations

from typing import Optional

import torch

from ..store.sharded_table import ShardedTable


class DeviceWideAndDeep:
    def __init__(self, num_slots: int, fields: int, emb_dim: int = 8, hidden: int = 64,
                 learning_rate: float = 0.05, pull_limit: int = 64, group=None, seed: int = 0,
                 device: Optional[int] = None):
        self.fields, self.emb_dim, self.lr, self.pull_limit = fields, emb_dim, learning_rate, pull_limit
        self.table = ShardedTable(num_slots, emb_dim + 1, group=group, device=device,
                                  init_range=(-0.05, 0.05), seed=seed)   # [embedding | wide weight]
        self.dev = self.table.cuda_device
        g = torch.Generator(device="cpu").manual_seed(seed)
        self.mlp = torch.nn.Sequential(torch.nn.Linear(fields * emb_dim, hidden), torch.nn.ReLU(),
                                       torch.nn.Linear(hidden, 1)).to(self.dev)
        for p in self.mlp.parameters():
            p.data = (torch.randn(p.shape, generator=g) * 0.1).to(self.dev)
        self.opt = torch.optim.SGD(self.mlp.parameters(), lr=learning_rate)

    def _pull(self, ids: torch.Tensor) -> torch.Tensor:
        # device-side pull limiter: one launch, at most `pull_limit` row pulls in flight
        return self.table.pull(ids.reshape(-1).contiguous(), pull_limit=self.pull_limit)

    def step(self, ids: torch.Tensor, labels: torch.Tensor) -> float:
        """ids: [B, fields] hashed feature slots, labels: [B] in {0,1}.  Returns the batch log-loss."""
        rows = self._pull(ids).requires_grad_(True)
        emb = rows[:, : self.emb_dim].reshape(ids.shape[0], -1)
        wide = rows[:, self.emb_dim].reshape(ids.shape[0], -1).sum(1)
        logit = self.mlp(emb).squeeze(1) + wide
        loss = torch.nn.functional.binary_cross_entropy_with_logits(logit, labels.float())
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()
        self.table.push(ids.reshape(-1), rows.grad.contiguous(), scale=-self.lr * ids.shape[0])
        return float(loss.detach())

    def predict(self, ids: torch.Tensor) -> torch.Tensor:
        with torch.no_grad():
            rows = self._pull(ids)
            emb = rows[:, : self.emb_dim].reshape(ids.shape[0], -1)
            wide = rows[:, self.emb_dim].reshape(ids.shape[0], -1).sum(1)
            return torch.sigmoid(self.mlp(emb).squeeze(1) + wide)

    def close(self):
        self.table.close()
