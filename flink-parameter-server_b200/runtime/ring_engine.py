"""Per-record ``WorkerLogic`` callbacks against a *device-resident, logic-bearing* parameter server.

``transform_rings`` runs the reference-shaped worker API (``onRecv`` / ``onPullRecv`` /
``ps.pull/push/output``) on the host while the server side -- the store, its registered update op
(add / assign / max / min) or its per-key lock logic (LockPSLogicA / LockPSLogicB) -- runs in the
persistent ``fps_server_loop`` kernel on the GPU that owns the shard; requests and answers travel
through the peer-memory rings and the pull limiter is the device credit counter
(ops/csrc/fps_rings.cu).  This is the device counterpart of ``transform(data, workerLogic, psLogic, ...)``
for the stores of M/server/*.scala; ids are ints, parameter values are float vectors.

Rank r is worker r and PS shard r.  The call returns when the local input is exhausted and every
pull of this worker has been answered (then a barrier, then the servers stop).
"""
from __future__ import annotations

import time
from typing import Any, Iterable, List, Optional

import torch

from ..api import Left, ParameterServerClient, Right
from ..parallel.rings import DeviceMessageServer, DeviceRingClient, RingFabric
from ..store.sharded_table import ShardedTable
from .stream import ResultStream


class _RingPSClient(ParameterServerClient):
    def __init__(self, outputs: List[Any]):
        self.pulls: List[int] = []
        self.pushes: List[tuple] = []
        self.outputs = outputs

    def pull(self, id) -> None:
        self.pulls.append(int(id))

    def push(self, id, deltaUpdate) -> None:
        self.pushes.append((int(id), deltaUpdate))

    def output(self, out) -> None:
        self.outputs.append(Left(out))


def transform_rings(local_data: Iterable[Any], workerLogic, table: ShardedTable, update: str = "add",
                    lock: Optional[str] = None, pull_limit: int = 1600, ring_capacity: int = 1024,
                    dump_model: bool = True, idle_timeout_s: float = 10.0) -> ResultStream:
    dev = table.cuda_device
    dim = table.dim
    rings = RingFabric(table.stride, capacity=ring_capacity, group=table.group, device=table.device)
    server = DeviceMessageServer(table, rings, update=update, lock=lock)
    client = DeviceRingClient(table, rings, pull_limit=pull_limit)
    outputs: List[Any] = []
    ps = _RingPSClient(outputs)
    server.start()
    if table.world > 1:
        import torch.distributed as dist

        dist.barrier(group=table.group)
    outstanding = 0

    def drain_requests():
        nonlocal outstanding
        if ps.pushes:
            ids = torch.tensor([i for i, _ in ps.pushes], dtype=torch.int64)
            vals = torch.stack([torch.as_tensor(v, dtype=torch.float32).reshape(-1)[:dim] for _, v in ps.pushes])
            client.push(ids, vals.to(dev))
            ps.pushes.clear()
        if ps.pulls:
            client.pull(torch.tensor(ps.pulls, dtype=torch.int64))
            outstanding += len(ps.pulls)
            ps.pulls.clear()

    def deliver(block: bool) -> None:
        nonlocal outstanding
        t0 = time.time()
        while outstanding > 0:
            ids, vals = client.collect(256)
            n = ids.numel()
            if n == 0:
                if not block:
                    return
                if time.time() - t0 > idle_timeout_s:
                    raise TimeoutError("device parameter server did not answer (a locked key was never pushed?)")
                time.sleep(0.0005)
                continue
            t0 = time.time()
            outstanding -= n
            vals = vals[:, :dim].cpu()
            for i, v in zip(ids.cpu().tolist(), vals):
                workerLogic.onPullRecv(i, v.clone(), ps)
            drain_requests()

    try:
        workerLogic.open()
        for rec in local_data:
            workerLogic.onRecv(rec, ps)
            drain_requests()
            deliver(block=False)
        deliver(block=True)
        workerLogic.close()
        if table.world > 1:
            import torch.distributed as dist

            dist.barrier(group=table.group)
    finally:
        server.stop()
    stats = server.stats()
    if dump_model:
        ids, vals = table.dump_local(only_touched=False) if server.touched is None else _touched(table, server)
        for i, v in zip(ids.cpu().tolist(), vals.cpu()):
            outputs.append(Right((i, v)))
    rings.close()
    out = ResultStream(outputs)
    out.server_stats = stats
    out.client_counters = client.counters()
    return out


def _touched(table: ShardedTable, server: DeviceMessageServer):
    slots = torch.arange(table.rows_per_shard, device=table.cuda_device)
    bits = (server.touched[slots >> 5] >> (slots & 31)) & 1
    ids = table.local_ids()
    sel = ((bits != 0) & (ids < table.num_ids)).nonzero(as_tuple=True)[0]
    return ids[sel], table.local[sel, : table.dim].clone()
