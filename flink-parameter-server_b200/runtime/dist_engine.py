"""Multi-process host tier: the same ``WorkerLogic`` / ``ParameterServerLogic`` callbacks with one
process per rank, messages exchanged through ``torch.distributed`` (gloo on CPU, NCCL works too).

Rank r hosts worker r and PS shard r (``workerParallelism == psParallelism == world_size``).  The
job advances in rounds; in every round each rank

  1. feeds up to ``records_per_round`` local input records to ``workerLogic.onRecv``,
  2. takes part in ONE exchange that carries both directions -- the worker->PS messages produced since the
     last round (bucketed by ``paramPartitioner``), the PS->worker answers produced in the last round
     (bucketed by ``wInPartition``) and the termination status,
  3. lets its PS shard handle the received requests (``onPullRecv`` / ``onPushRecv``; the answers travel in
     the next round) and delivers the received answers (``onPullRecv``; its pushes / pulls travel in the next
     round).

Message lists keep their order, so delivery is FIFO per (producer, consumer) pair like Flink's
channels.  The job ends when every rank is out of input and a full round moved no message (the
distributed form of the local engine's quiescence detection; the reference uses a timeout,
FPS:480).  This is the plumbing path of BASELINE.json config 1 (CPU, world_size=2).
"""
from __future__ import annotations

from typing import Any, Callable, Iterable, List, Optional

import torch.distributed as dist

from ..api import Left, ParameterServer, ParameterServerClient, Right, RuntimeContext
from ..protocol.senders import (SimplePSReceiver, SimplePSSender, SimpleWorkerReceiver,
                                SimpleWorkerSender)
from .stream import ResultStream
from ..parallel.partitioner import stable_hash
from .transform import default_param_partitioner, default_worker_partitioner


def _exchange(payload: Any, group) -> List[Any]:
    """One pickled all-gather per round: every rank's ``payload`` (its per-destination buckets + status)."""
    world = dist.get_world_size(group)
    gathered: List[Any] = [None] * world
    dist.all_gather_object(gathered, payload, group=group)
    return gathered


def transform_distributed(local_data: Iterable[Any], workerLogic, psLogic,
                          paramPartitioner: Optional[Callable[[Any], int]] = None,
                          wInPartition: Optional[Callable[[Any], int]] = None, group=None,
                          records_per_round: int = 1024, gather_results: bool = True) -> ResultStream:
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    part = paramPartitioner or default_param_partitioner(world)
    wpart = wInPartition or default_worker_partitioner(world)
    w_send, w_recv = SimpleWorkerSender(), SimpleWorkerReceiver()
    p_send, p_recv = SimplePSSender(), SimplePSReceiver()
    results: List[Any] = []
    to_ps: List[List[Any]] = [[] for _ in range(world)]
    to_worker: List[List[Any]] = [[] for _ in range(world)]

    class Client(ParameterServerClient):
        def pull(self, id):
            w_send.onPull(id, lambda m: to_ps[int(part(m)) % world].append(m), rank)

        def push(self, id, deltaUpdate):
            w_send.onPush(id, deltaUpdate, lambda m: to_ps[int(part(m)) % world].append(m), rank)

        def output(self, out):
            results.append(Left(out))

    class Server(ParameterServer):
        def answerPull(self, id, value, workerPartitionIndex):
            def collect(m):
                d = int(wpart(m))
                if not 0 <= d < world:
                    raise RuntimeError("Pull answer key should be the partition ID itself!")
                to_worker[d].append(m)

            p_send.onPullAnswer(id, value, workerPartitionIndex, collect)

        def output(self, out):
            results.append(Right(out))

    # default routing (hash of the id / worker index): plain tuples on the wire instead of message objects --
    # (0, id, worker) = pull, (1, id, delta) = push, (id, value) = answer; cheaper to build, route and pickle
    fast = paramPartitioner is None and wInPartition is None

    class FastClient(ParameterServerClient):
        def pull(self, id):
            to_ps[stable_hash(id) % world].append((0, id, rank))

        def push(self, id, deltaUpdate):
            to_ps[stable_hash(id) % world].append((1, id, deltaUpdate))

        def output(self, out):
            results.append(Left(out))

    class FastServer(ParameterServer):
        def answerPull(self, id, value, workerPartitionIndex):
            d = int(workerPartitionIndex)
            if not 0 <= d < world:
                raise RuntimeError("Pull answer key should be the partition ID itself!")
            to_worker[d].append((id, value))

        def output(self, out):
            results.append(Right(out))

    client, server = (FastClient(), FastServer()) if fast else (Client(), Server())
    workerLogic.open()
    psLogic.open({}, RuntimeContext(rank, world))
    it = iter(local_data)
    exhausted = False
    on_pull = lambda id, widx: psLogic.onPullRecv(id, widx, server)
    on_push = lambda id, delta: psLogic.onPushRecv(id, delta, server)
    on_answer = lambda a: workerLogic.onPullRecv(a.paramId, a.param, client)
    while True:
        fed = 0
        while not exhausted and fed < records_per_round:
            try:
                workerLogic.onRecv(next(it), client)
                fed += 1
            except StopIteration:
                exhausted = True
        out_ps, to_ps[:] = list(to_ps), [[] for _ in range(world)]
        out_w, to_worker[:] = list(to_worker), [[] for _ in range(world)]
        n_moved = sum(len(b) for b in out_ps) + sum(len(b) for b in out_w)
        gathered = _exchange((out_ps, out_w, exhausted, n_moved), group)
        for src in range(world):                      # FIFO per (producer, consumer) pair
            if fast:
                for m in gathered[src][0][rank]:
                    if m[0] == 0:
                        psLogic.onPullRecv(m[1], m[2], server)
                    else:
                        psLogic.onPushRecv(m[1], m[2], server)
            else:
                for m in gathered[src][0][rank]:
                    p_recv.onWorkerMsg(m, on_pull, on_push)
        for src in range(world):
            if fast:
                for m in gathered[src][1][rank]:
                    workerLogic.onPullRecv(m[0], m[1], client)
            else:
                for m in gathered[src][1][rank]:
                    w_recv.onPullAnswerRecv(m, on_answer)
        # every rank sees the same statuses: stop when all inputs are exhausted and the round carried nothing
        # (then nothing was delivered anywhere, so no rank holds a message produced by this round either)
        if all(g[2] and g[3] == 0 for g in gathered):
            break
    workerLogic.close()
    psLogic.close(server)
    if gather_results:
        allr = [None] * world
        dist.all_gather_object(allr, results, group=group)
        results = [x for part_ in allr for x in part_]
    return ResultStream(results)
