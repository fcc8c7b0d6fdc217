"""Server-side stores (layer L4): the seven ``ParameterServerLogic`` implementations.

Host-tier (generic Python values, any hashable id) versions of M/server/*.scala.  The device
tier implements the same semantics on dense HBM shards (store/sharded_table.py: lazy init by
Philox, additive update by ``red.add``, touched bitmap, dump at close) and, for lock stores, in
the persistent ``fps_server_loop`` kernel (ops/csrc/fps_rings.cu).
"""
from __future__ import annotations

import math
from collections import deque
from typing import Any, Callable, Deque, Dict, List, Optional, Tuple

from ..api import (LooseParameterServerLogic, ParameterServer, ParameterServerLogic, RuntimeContext)


class SimplePSLogic(ParameterServerLogic):
    """HashMap store: init on first pull, update on push, emit ``(id, value)`` on EVERY push.

    Push to an unseen id stores the delta as the value (SimplePSLogic.scala:7-26).
    """

    def __init__(self, paramInit: Callable[[Any], Any], paramUpdate: Callable[[Any, Any], Any]):
        self.init = paramInit
        self.update = paramUpdate
        self.params: Dict[Any, Any] = {}

    def onPullRecv(self, id, workerPartitionIndex, ps):
        if id not in self.params:
            self.params[id] = self.init(id)
        ps.answerPull(id, self.params[id], workerPartitionIndex)

    def onPushRecv(self, id, deltaUpdate, ps):
        c = self.update(self.params[id], deltaUpdate) if id in self.params else deltaUpdate
        self.params[id] = c
        ps.output((id, c))


class SimplePSLogicWithClose(SimplePSLogic):
    """Same store, but the model is emitted only at ``close()`` (SimplePSLogicWithClose.scala:7-32)."""

    def onPushRecv(self, id, deltaUpdate, ps):
        self.params[id] = (self.update(self.params[id], deltaUpdate) if id in self.params
                           else deltaUpdate)

    def close(self, ps):
        for id, c in self.params.items():
            ps.output((id, c))


class LooseSimplePSLogic(LooseParameterServerLogic):
    """PullP != PushP.  Push to an unseen id stores ``init(id)`` -- the first delta is dropped
    (LooseSimplePSLogic.scala:21-24; kept for parity, see ``drop_first_delta``)."""

    def __init__(self, paramInit, paramUpdate, drop_first_delta: bool = True):
        self.init = paramInit
        self.update = paramUpdate
        self.drop_first_delta = drop_first_delta
        self.params: Dict[Any, Any] = {}

    def onPullRecv(self, id, workerPartitionIndex, ps):
        if id not in self.params:
            self.params[id] = self.init(id)
        ps.answerPull(id, self.params[id], workerPartitionIndex)

    def onPushRecv(self, id, deltaUpdate, ps):
        if id in self.params:
            c = self.update(self.params[id], deltaUpdate)
        elif self.drop_first_delta:
            c = self.init(id)
        else:
            c = self.update(self.init(id), deltaUpdate)
        self.params[id] = c
        ps.output((id, c))


class LooseSimplePSLogicWithClose(LooseParameterServerLogic):
    """Loose store with ``store: PushP => PullP`` for unseen ids; dump at close
    (LooseSimplePSLogicWithClose.scala:7-34)."""

    def __init__(self, paramInit, paramUpdate, store: Callable[[Any], Any]):
        self.init = paramInit
        self.update = paramUpdate
        self.store = store
        self.params: Dict[Any, Any] = {}

    def onPullRecv(self, id, workerPartitionIndex, ps):
        if id not in self.params:
            self.params[id] = self.init(id)
        ps.answerPull(id, self.params[id], workerPartitionIndex)

    def onPushRecv(self, id, deltaUpdate, ps):
        self.params[id] = (self.update(self.params[id], deltaUpdate) if id in self.params
                           else self.store(deltaUpdate))

    def close(self, ps):
        for id, c in self.params.items():
            ps.output((id, c))


class RangePSLogicWithClose(ParameterServerLogic):
    """Dense array over this shard's contiguous id range (RangePSLogicWithClose.scala:7-63).

    ``open`` computes ``div = ceil(featureCount / n)``, ``startIndex = subtask * div``; the last
    shard holds the remainder.  Only initialised slots are dumped at close.
    """

    def __init__(self, featureCount: int, paramInit, paramUpdate):
        self.featureCount = int(featureCount)
        self.init = paramInit
        self.update = paramUpdate
        self.startIndex = 0
        self.params: List[Optional[Any]] = []
        self._set: List[bool] = []

    def open(self, parameters, runtimeContext: RuntimeContext):
        n = runtimeContext.getNumberOfParallelSubtasks()
        idx = runtimeContext.getIndexOfThisSubtask()
        div = int(math.ceil(self.featureCount / n))
        mod = self.featureCount - (n - 1) * div
        size = mod if (mod != 0 and idx + 1 == n) else div
        size = max(size, 0)
        self.params = [None] * size
        self._set = [False] * size
        self.startIndex = idx * div

    def onPullRecv(self, id, workerPartitionIndex, ps):
        i = id - self.startIndex
        if not self._set[i]:
            self.params[i] = self.init(id)
            self._set[i] = True
        ps.answerPull(id, self.params[i], workerPartitionIndex)

    def onPushRecv(self, id, deltaUpdate, ps):
        i = id - self.startIndex
        self.params[i] = self.update(self.params[i], deltaUpdate) if self._set[i] else deltaUpdate
        self._set[i] = True

    def close(self, ps):
        for i, ok in enumerate(self._set):
            if ok:
                ps.output((self.startIndex + i, self.params[i]))


class _LockEntry:
    __slots__ = ("locked", "value", "queue")

    def __init__(self, value):
        self.locked = False
        self.value = value
        self.queue: Deque[int] = deque()


class LockPSLogicA(ParameterServerLogic):
    """Per-key lock: a pull takes the lock, later pulls queue, a push releases it or hands the
    fresh value to the queue head and stays locked; push to an unknown id raises
    (LockPSLogicA.scala:13-46).  Duplicate waiters are kept."""

    dedup_waiters = False

    def __init__(self, init, update):
        self.init = init
        self.update = update
        self.params: Dict[Any, _LockEntry] = {}

    def onPullRecv(self, id, workerPartitionIndex, ps):
        e = self.params.get(id)
        if e is None:
            e = _LockEntry(self.init(id))
            self.params[id] = e
        if not e.locked:
            ps.answerPull(id, e.value, workerPartitionIndex)
            e.locked = True
        elif not (self.dedup_waiters and workerPartitionIndex in e.queue):
            e.queue.append(workerPartitionIndex)

    def onPushRecv(self, id, deltaUpdate, ps):
        e = self.params.get(id)
        if e is None:
            raise RuntimeError("Not existed model was not able to update by any delta.")
        c = self.update(e.value, deltaUpdate)
        e.value = c
        if not e.queue:
            e.locked = False
        else:
            ps.answerPull(id, c, e.queue.popleft())
            e.locked = True
        ps.output((id, c))

    # introspection used by tests (mirrors `params(id) = (locked, value, queue)`)
    def state(self, id) -> Tuple[bool, Any, List[int]]:
        e = self.params[id]
        return e.locked, e.value, list(e.queue)


class LockPSLogicB(LockPSLogicA):
    """Same, but a worker index waits at most once per key (LockPSLogicB.scala:15-50)."""

    dedup_waiters = True
