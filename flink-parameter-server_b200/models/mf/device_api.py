"""``psOnlineMF(..., backend="device")`` / ``psOfflineMF(..., backend="device")``: reference-shaped
entry points over :class:`DeviceOnlineMF`.

``src`` is an iterable of ``Rating`` records (any size; micro-batched on the host, packed64, pinned)
or of ready ``(users, items, ratings)`` host tensors.  In a multi-rank job every rank passes its own
partition of the stream (users with ``user % world == rank``), like Flink's parallel sources behind
``partitionCustom(user % n)``.  The result stream holds ``Left((userId, vector))`` for the local
users touched and ``Right((itemId, vector))`` for the local shard's items (model dump at close).
"""
from __future__ import annotations

from typing import Iterable, Iterator, List, Optional, Sequence

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


def ps_online_mf_device(src, numFactors=10, rangeMin=-0.01, rangeMax=0.01, learningRate=0.01,
                        negativeSampleRate=0, pullLimit=0, seed=0, plain_residual=False,
                        numUsers: Optional[int] = None, numItems: Optional[int] = None,
                        batch_size: int = 1 << 16, group=None, epochs: int = 1,
                        userMemory: int = 128) -> ResultStream:
    recs = None
    if numUsers is None or numItems is None:
        recs = list(src.collect() if hasattr(src, "collect") else src)
        numUsers = 1 + max(r.user for r in recs)
        numItems = 1 + max(r.item for r in recs)
        src = recs
    # the reference's default pullLimit (1600) is a JVM-queue bound; on the device tier 0 means
    # "as many row slots in flight as the GPU holds" and an explicit value bounds the rows in flight
    model = DeviceOnlineMF(numUsers, numItems, numFactors, rangeMin, rangeMax, learningRate,
                           negativeSampleRate, pull_limit=pullLimit if pullLimit and pullLimit != 1600 else 0,
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
