"""Generic tensor tier: run user ``BatchedWorkerLogic`` callbacks against a device-resident sharded
table ("drives workerParallelism WorkerLogic loops on CUDA streams fed by a DataStream-shaped
iterator").

``transform_device(batches, workerLogic, table, ...)`` is the device analogue of
``FlinkParameterServer.transform``: every rank is worker ``rank`` (and PS shard ``rank``); each
micro-batch is handed to ``logic.onRecvBatch(batch, ps)``; ``ps.pull(ids)`` runs the one-sided gather
kernel on the worker's stream and immediately calls ``logic.onPullRecvBatch(ids, values, ps)``
(stream ordered, no host synchronisation); ``ps.push(ids, deltas)`` runs the ``red.add`` kernel
(= additive ``paramUpdate`` in the owner's memory system).  ``pull_limit`` bounds the number of rows
gathered per kernel (chunked pulls).  With ``worker_streams > 1`` several logic clones run
round-robin on their own CUDA streams, like ``workerParallelism > #GPUs`` in the reference.

Outputs: ``ps.output(x)`` -> ``Left(x)``; at the end the local shard is dumped as ``Right((id, row))``
for touched ids (``*WithClose`` semantics) when ``dump_model=True``.
"""
from __future__ import annotations

from typing import Any, Iterable, List

import torch

from ..api import BatchedParameterServerClient, BatchedWorkerLogic, Left, Right
from ..store.sharded_table import ShardedTable
from .local_engine import clone_logic
from .stream import ResultStream


class DeviceParameterServerClient(BatchedParameterServerClient):
    def __init__(self, table: ShardedTable, logic: BatchedWorkerLogic, outputs: List[Any],
                 pull_limit: int = 0):
        self.table, self.logic, self.outputs, self.pull_limit = table, logic, outputs, int(pull_limit)
        self.pulled_rows = 0
        self.pushed_rows = 0

    def pull_now(self, ids: torch.Tensor) -> torch.Tensor:
        self.pulled_rows += ids.numel()
        return self.table.pull(ids, pull_limit=self.pull_limit)

    def pull(self, ids: torch.Tensor) -> None:
        lim = self.pull_limit if self.pull_limit > 0 else ids.numel()
        for a in range(0, ids.numel(), max(1, lim)):      # credit-sized chunks, FIFO order
            chunk = ids[a:a + lim]
            self.logic.onPullRecvBatch(chunk, self.pull_now(chunk), self)

    def push(self, ids: torch.Tensor, deltaUpdate: torch.Tensor) -> None:
        self.pushed_rows += ids.numel()
        self.table.push(ids, deltaUpdate.contiguous())

    def output(self, out: Any) -> None:
        self.outputs.append(Left(out))


def transform_batched(batches: Iterable[Any], workerLogic: BatchedWorkerLogic, paramInit="zeros",
                      paramUpdate="add", workerParallelism: int = 1, psParallelism: int = 1,
                      iterationWaitTime: float = 0, *, num_ids: int, dim: int, pull_limit: int = 0,
                      paramPartitioner=None, partition: str = "hash", seed: int = 0,
                      dump_model: bool = True, table: ShardedTable = None, model=None) -> ResultStream:
    """``transform(batches, BatchedWorkerLogic, paramInit, paramUpdate, workerParallelism,
    psParallelism, ..., backend="device")`` -- the front door of the tensor tier (FPS:64-80 shape).

    ``paramInit``: ``"zeros"``, a float constant, or ``("uniform", lo, hi)`` (Philox by id);
    ``paramUpdate``: ``"add"`` (the additive update fused into the push).  ``psParallelism`` shards:
    one per rank in a multi-process job (``psParallelism <= world``), otherwise ``psParallelism``
    logical shards on this process's GPUs (:class:`MultiShardTable`); ``paramPartitioner(id) -> shard``
    (any function) becomes a device ``(owner, slot)`` lookup table.  ``workerParallelism`` worker
    loops run on their own CUDA streams (per process: ``ceil(wP / world)``)."""
    import torch.distributed as dist

    if paramUpdate not in ("add", "sum"):
        raise ValueError('the tensor tier fuses an additive paramUpdate into the push: paramUpdate="add"')
    if isinstance(paramInit, tuple) and paramInit and paramInit[0] == "uniform":
        init, rng, const = "uniform", (float(paramInit[1]), float(paramInit[2])), None
    elif paramInit == "zeros" or paramInit == 0:
        init, rng, const = "zeros", (0.0, 0.0), None
    elif isinstance(paramInit, (int, float)):
        init, rng, const = "zeros", (0.0, 0.0), float(paramInit)
    else:
        raise ValueError('paramInit must be "zeros", a constant or ("uniform", lo, hi) on the tensor tier')
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    own = table is None
    if own:
        if world > 1:
            if paramPartitioner is not None:
                raise NotImplementedError("custom partitioners need one LUT copy per process: single-process only")
            table = ShardedTable(num_ids, dim, partition=partition, init=init, init_range=rng, seed=seed,
                                 track_touched=True, num_shards=psParallelism)
        else:
            from ..store.multi_shard import MultiShardTable

            table = MultiShardTable(num_ids, dim, psParallelism, partition=partition,
                                    partitioner=paramPartitioner, init=init, init_range=rng, seed=seed,
                                    track_touched=True)
        if const is not None:
            for t in getattr(table, "shards", [table.local] if table.owns_shard else []):
                t[:, :dim] = const
            table.barrier()
    if model is not None:        # transformWithModelLoad: every entry reaches its shard before training
        entries = list(model.collect() if hasattr(model, "collect") else model)
        if entries:
            dev = table.cuda_device
            ids = torch.tensor([int(k) for k, _ in entries], dtype=torch.int64, device=dev)
            vals = torch.stack([torch.as_tensor(v, dtype=torch.float32).reshape(-1) for _, v in entries]).to(dev)
            table.load(ids, vals)
        table.barrier()
    streams = -(-int(workerParallelism) // world)
    out = transform_device(batches, workerLogic, table, pull_limit=pull_limit, worker_streams=streams,
                           dump_model=dump_model)
    out.table = table
    return out


def transform_device(batches: Iterable[Any], workerLogic: BatchedWorkerLogic, table: ShardedTable,
                     pull_limit: int = 0, worker_streams: int = 1, dump_model: bool = True) -> ResultStream:
    dev = table.cuda_device
    outputs: List[Any] = []
    n = max(1, int(worker_streams))
    logics = [workerLogic if n == 1 else clone_logic(workerLogic) for _ in range(n)]
    streams = [torch.cuda.current_stream(dev)] if n == 1 else [torch.cuda.Stream(device=dev) for _ in range(n)]
    clients = [DeviceParameterServerClient(table, lg, outputs, pull_limit) for lg in logics]
    for lg in logics:
        lg.open()
    main = torch.cuda.current_stream(dev)
    for s in streams:
        s.wait_stream(main)
    for k, batch in enumerate(batches):
        w = k % n
        with torch.cuda.stream(streams[w]):
            logics[w].onRecvBatch(batch, clients[w])
    for s in streams:
        main.wait_stream(s)
    for lg in logics:
        lg.close()
    table.barrier()
    table.check_finite()
    if dump_model:
        ids, vals = table.dump_local()
        for i, v in zip(ids.cpu().tolist(), vals.cpu()):
            outputs.append(Right((i, v)))
    out = ResultStream(outputs)
    out.clients = clients
    return out
