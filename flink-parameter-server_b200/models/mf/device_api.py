"""``psOnlineMF(..., backend="device")`` / ``psOfflineMF(..., backend="device")``: reference-shaped
entry points over :class:`DeviceOnlineMF`.

``src`` is an iterable of ``Rating`` records (any size; micro-batched on the host, packed64, pinned)
or of ready ``(users, items, ratings)`` host tensors.  In a multi-rank job every rank passes its own
partition of the stream (users with ``user % world == rank``), like Flink's parallel sources behind
``partitionCustom(user % n)``.  The result stream holds ``Left((userId, vector))`` for the local
users touched and ``Right((itemId, vector))`` for the local shard's items (model dump at close).
"""
from __future__ import annotations

from typing import Iterator, List, Optional, Sequence

import torch

from ...api import Left, Right
from ...ops import native
from ...runtime.stream import ResultStream, as_stream
from .common import Rating
from .device import ERR_PLAIN, ERR_SIGMOID, DeviceOnlineMF


def _batches(src, batch_size: int) -> Iterator[Sequence[torch.Tensor]]:
    buf: List[Rating] = []
    for rec in (src.collect() if hasattr(src, "collect") else src):
        if isinstance(rec, (tuple, list)) and len(rec) in (1, 3) and torch.is_tensor(rec[0]):
            yield tuple(t if t.is_pinned() or not torch.cuda.is_available() else t.pin_memory() for t in rec)
            continue
        buf.append(rec)
        if len(buf) >= batch_size:
            yield _pack(buf)
            buf = []
    if buf:
        yield _pack(buf)


def _pack(buf: List[Rating]):
    u = torch.tensor([r.user for r in buf], dtype=torch.int32)
    i = torch.tensor([r.item for r in buf], dtype=torch.int32)
    r = torch.tensor([r.rating for r in buf], dtype=torch.float32)
    return (u.pin_memory(), i.pin_memory(), r.pin_memory())


def _result(model: DeviceOnlineMF, seen_users: Optional[set]) -> ResultStream:
    out = []
    uid, uvec = model.user_vectors()
    for i, v in zip(uid.cpu().tolist(), uvec.cpu().double().numpy()):
        if seen_users is None or i in seen_users:
            out.append(Left((i, v)))
    iid, ivec = model.item_vectors()
    for i, v in zip(iid.cpu().tolist(), ivec.cpu().double().numpy()):
        out.append(Right((i, v)))
    rs = ResultStream(out)
    rs.model = model
    return rs


def _agree_max(num_users: int, num_items: int, group):
    """Sizes derived from a rank's LOCAL partition differ between ranks; every rank must build the same
    table geometry (peer row addresses are computed from it), so take the maximum over the job."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return num_users, num_items
    sizes = [None] * dist.get_world_size(group)
    dist.all_gather_object(sizes, (int(num_users), int(num_items)), group=group)
    return max(s[0] for s in sizes), max(s[1] for s in sizes)


def ps_online_mf_device(src, numFactors=10, rangeMin=-0.01, rangeMax=0.01, learningRate=0.01,
                        negativeSampleRate=0, pullLimit=0, seed=0, plain_residual=False,
                        numUsers: Optional[int] = None, numItems: Optional[int] = None,
                        batch_size: int = 1 << 16, group=None, epochs: int = 1,
                        userMemory: int = 128) -> ResultStream:
    recs = None
    if numUsers is None or numItems is None:
        recs = list(src.collect() if hasattr(src, "collect") else src)
        numUsers = 1 + max((r.user for r in recs), default=0)
        numItems = 1 + max((r.item for r in recs), default=0)
        src = recs
        numUsers, numItems = _agree_max(numUsers, numItems, group)
    # pullLimit: None or 0 = as many row slots in flight as the GPU holds; any positive value bounds the
    # rows in flight on the device (the reference's 1600 default is a JVM-queue bound -- pass it
    # explicitly if that is what is wanted)
    model = DeviceOnlineMF(numUsers, numItems, numFactors, rangeMin, rangeMax, learningRate,
                           negativeSampleRate, pull_limit=int(pullLimit or 0),
                           group=group, seed=seed, err_mode=ERR_PLAIN if plain_residual else ERR_SIGMOID,
                           track_touched=True,
                           user_memory=min(int(userMemory), 256) if negativeSampleRate > 0 else 0)
    seen = set()
    data = list(_batches(src, batch_size))
    for b in data:
        if len(b) == 3:
            seen.update(b[0].tolist())
    for _ in range(max(1, epochs)):
        for _loss in model.fit_stream(iter(data)):
            pass
    model.check_finite()
    model.barrier()
    return _result(model, seen if seen else None)


def ps_offline_mf_device(src, iterations=10, **kw) -> ResultStream:
    """Multi-epoch variant: the finite stream is buffered once and replayed ``iterations`` times
    (PSOfflineMatrixFactorizationWorker.scala:97-128)."""
    return ps_online_mf_device(src, epochs=iterations, **kw)


# ---------------------------------------------------------------------------------------------------
# psTopKGenerator(..., backend="device") / psOnlineLearnerAndGenerator(..., backend="device")
# ---------------------------------------------------------------------------------------------------
class _ListWithModel(list):
    """A plain result list (like the host tier returns) that also carries the trained model."""

    model = None


def _seen_filter(rows, K: int, memory: int):
    """The sequential part of ``CollectTopKFromEachWorker`` (utils/CollectTopKFromEachWorker.scala:41-56):
    drop items in the user's recent set (bounded by ``memory``; -1 = unbounded), keep ``K``, then remember
    the rated item.  ``rows``: iterable of ``(user, item, ts, [(score, itemId), ...])`` in stream order."""
    from collections import deque

    seen, order, out = {}, {}, []
    for user, item, ts, cand in rows:
        s = seen.setdefault(user, set())
        out.append((user, item, ts, [c for c in cand if c[1] not in s][:K]))
        s.add(item)
        q = order.setdefault(user, deque())
        q.append(item)
        if memory > -1 and len(q) > memory:
            s.discard(q.popleft())
    return out


def ps_topk_generator_device(src, model, K: int = 100, workerK: int = 75, userMemory: int = 0,
                             batch_size: int = 4096, group=None, sort_by_length: bool = True):
    """Top-K serving over a pre-trained model on the device tier (capability of ``psTopKGenerator``,
    PSTopKGenerator.scala:47-107): the user vectors of ``model`` are loaded into a sharded PS table, the
    item vectors stay with this worker, every query is scored on the tcgen05 kernel (length-sorted item
    table = the LEMP LENGTH bound) and the per-worker lists are merged.  ``model`` has the reference's
    orientation: ``Left((itemId, (len, vec)))`` / ``Right((userId, (len, vec)))``.  In a multi-rank job
    every rank passes its own part of the model and the *same* query stream.  Unknown users get an empty
    list (the reference's ``invalidParam``).  Returns ``[(itemId, timestamp, [(score, itemId)])]``."""
    import numpy as np
    import torch.distributed as dist

    from ...store.sharded_table import ShardedTable
    from .device_topk import DistributedTopK

    dev = torch.device("cuda", torch.cuda.current_device())
    entries = list(model.collect() if hasattr(model, "collect") else model)
    items = [(e.value[0], np.asarray(e.value[1][1], dtype=np.float32)) for e in entries if e.is_left]
    users = [(e.value[0], np.asarray(e.value[1][1], dtype=np.float32)) for e in entries if not e.is_left]
    k = len(items[0][1]) if items else len(users[0][1])
    n_users = 1 + max([u for u, _ in users], default=0)
    ready = dist.is_available() and dist.is_initialized()
    if ready:
        t = torch.tensor([n_users, k], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        n_users, k = int(t[0]), int(t[1])
    # one extra column flags "this user was loaded": item rows carry 0 there, so scores are unchanged
    table = ShardedTable(n_users, k + 1, group=group, init="zeros")
    if users:
        uid = torch.tensor([u for u, _ in users], dtype=torch.int64, device=dev)
        uvec = torch.zeros((len(users), k + 1), dtype=torch.float32, device=dev)
        uvec[:, :k] = torch.from_numpy(np.stack([v for _, v in users])).to(dev)
        uvec[:, k] = 1.0
        table.load(uid, uvec)
    table.barrier()
    local = torch.zeros((max(len(items), 1), table.stride), dtype=torch.float32, device=dev)
    if items:
        local[: len(items), :k] = torch.from_numpy(np.stack([v for _, v in items])).to(dev)
        item_ids = torch.tensor([i for i, _ in items], dtype=torch.int64, device=dev)
    else:
        item_ids = torch.full((1,), -1, dtype=torch.int64, device=dev)
    serving = DistributedTopK(table, local, item_ids, group=group)
    if sort_by_length and len(items) >= 2048:
        from .device_topk import DeviceTopK

        serving.local = DeviceTopK(local, sort_by_length=True)
    ratings = list(src.collect() if hasattr(src, "collect") else src)
    want = K + (min(userMemory, 4 * K) if userMemory >= 0 else 4 * K)   # room for the seen-item filter
    rows = []
    for a in range(0, len(ratings), batch_size):
        chunk = ratings[a:a + batch_size]
        q = torch.tensor([min(max(r.user, 0), n_users - 1) for r in chunk], dtype=torch.int64, device=dev)
        known = (table.pull(q)[:, k] == 1.0).cpu().tolist()
        sc, ids = serving.topk(q, want, workerK=max(workerK, want))
        sc, ids = sc.cpu().tolist(), ids.cpu().tolist()
        for j, r in enumerate(chunk):
            ok = known[j] and 0 <= r.user < n_users
            cand = [(s, i) for s, i in zip(sc[j], ids[j]) if i >= 0 and s > -1.0e38] if ok else []
            rows.append((r.user, r.item, r.getEventTime(), cand))
    table.close()
    return [(item, ts, topk) for (_u, item, ts, topk) in _seen_filter(rows, K, userMemory)]


def ps_online_learner_and_generator_device(src, numFactors=10, rangeMin=-0.001, rangeMax=0.001,
                                           learningRate=0.01, negativeSampleRate=0, userMemory=65535,
                                           K=100, pullLimit=0, seed=0, plain_residual=False,
                                           numUsers: Optional[int] = None, numItems: Optional[int] = None,
                                           batch_size: int = 4096, group=None):
    """Online MF plus a top-K list for every incoming rating, computed BEFORE the model sees that rating
    (prequential evaluation; capability of ``psOnlineLearnerAndGenerator``,
    PSOnlineMatrixFactorizationAndTopKGenerator.scala:51-101) on the device tier: per micro-batch
    (1) score the batch's users against the item table with the tensor-core kernel, (2) train on the
    batch with the fused kernel.  Prequential at micro-batch granularity (``batch_size``).  Single rank.
    Returns ``[(userId, itemId, timestamp, [(score, itemId)])]`` and leaves the model in ``.model``."""
    import torch.distributed as dist

    from .device_topk import DeviceTopK

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        raise NotImplementedError("the device learner+generator runs on one rank; use DistributedTopK "
                                  "with DeviceOnlineMF for multi-rank serving")
    recs = list(src.collect() if hasattr(src, "collect") else src)
    if numUsers is None:
        numUsers = 1 + max(r.user for r in recs)
    if numItems is None:
        numItems = 1 + max(r.item for r in recs)
    model = DeviceOnlineMF(numUsers, numItems, numFactors, rangeMin, rangeMax, learningRate,
                           negativeSampleRate, pull_limit=int(pullLimit or 0),
                           group=group, seed=seed, err_mode=ERR_PLAIN if plain_residual else ERR_SIGMOID,
                           item_cache=False, user_memory=min(max(userMemory, 0), 256) if negativeSampleRate else 0)
    dev = model.cuda_device
    want = K + (min(userMemory, 4 * K) if userMemory >= 0 else K)
    rows = []
    for a in range(0, len(recs), batch_size):
        chunk = recs[a:a + batch_size]
        u = torch.tensor([r.user for r in chunk], dtype=torch.int32, device=dev)
        i = torch.tensor([r.item for r in chunk], dtype=torch.int32, device=dev)
        rt = torch.tensor([r.rating for r in chunk], dtype=torch.float32, device=dev)
        q = model.users[u.long()].contiguous()                       # world == 1: slot == user id
        sc, ids = DeviceTopK(model.items.local[:numItems]).topk(want, q_local=q)
        sc, ids = sc.cpu().tolist(), ids.cpu().tolist()
        for j, r in enumerate(chunk):
            rows.append((r.user, r.item, r.getEventTime(), list(zip(sc[j], ids[j]))))
        model.step(u, i, rt)
    model.check_finite()
    out = _ListWithModel(_seen_filter(rows, K, userMemory))
    out.model = model
    return out
