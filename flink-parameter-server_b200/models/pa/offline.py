"""Legacy multi-epoch train-then-predict passive-aggressive pipeline --
``paBinaryClassificationOffline`` and its (name-only different) ``paMultiClassificationOffline``.

Reference: M/passive/aggressive/classification/binary/PABinaryClassificationOffline.scala:47-387
(``multi/PAMultiClassificationOffline.scala`` is a byte-identical copy) with the legacy
map-based ``PassiveAggressiveFilter`` (algorithm/binary/PassiveAggressiveFilter.scala:7-49, C54).

Pipeline: training vectors ``(LegacySparseVector, label in {+1,-1})`` are shuffled over the workers
and buffered until the end of the training input; a background thread then issues the pulls for
``iterations`` epochs behind a *blocking* pull limit (``ReentrantLock``/``Condition`` in the
reference -> :func:`addBlockingPullLimiter`); once training pulls are all answered the test vectors
are predicted; predictions are emitted as worker outputs ``(vector, predictedLabel)`` (the reference
logs them at ``close()``), the model as PS outputs ``(featureId, weight)``.
"""
from __future__ import annotations

import random
import threading
from collections import deque
from typing import Dict, List, Optional, Tuple

from ...api import CtorFork, WorkerLogic
from ...limiter import addBlockingPullLimiter
from ...runtime.stream import DataStream, as_stream
from ...runtime.transform import transform
from ...server.logics import SimplePSLogicWithClose
from ...utils.eof import EOF
from .sparse import LegacySparseVector


class PassiveAggressiveFilter:
    """Legacy binary PA on map-based vectors: ``delta(data, model, label)`` / ``predict``."""

    def __init__(self, C: float = 0):
        self.Const = C

    @staticmethod
    def buildPAF():
        return PassiveAggressiveFilterImp()

    @staticmethod
    def buildPAFI(Con):
        return PassiveAggressiveFilterImpI(Con)

    @staticmethod
    def buildPAFII(Con):
        return PassiveAggressiveFilterImpII(Con)

    def getTau(self, data: LegacySparseVector, l: float) -> float:
        raise NotImplementedError

    def delta(self, data: LegacySparseVector, model: Dict[int, float], label: int) -> Dict[int, float]:
        assert label in (1, -1)
        assert set(data.getIndexes()) == set(model.keys())
        l = max(0.0, 1 - label * sum(model[k] * v for k, v in data.getValues().items()))
        mult = self.getTau(data, l) * label
        return {k: v * mult for k, v in data.getValues().items()}

    def predict(self, data: LegacySparseVector, model: Dict[int, float]) -> int:
        s = sum(model.get(k, 0.0) * v for k, v in data.getValues().items())
        return (s > 0) - (s < 0)

    @staticmethod
    def quotient(data: LegacySparseVector, l: float, denominatorConst: float) -> float:
        n2 = sum(v * v for v in data.getValues().values())
        return l / n2 if denominatorConst == 0 else l / (n2 + denominatorConst)


class PassiveAggressiveFilterImp(PassiveAggressiveFilter):
    def getTau(self, data, l):
        return self.quotient(data, l, 0)


class PassiveAggressiveFilterImpI(PassiveAggressiveFilter):
    def getTau(self, data, l):
        return min(self.Const, self.quotient(data, l, 0))


class PassiveAggressiveFilterImpII(PassiveAggressiveFilter):
    def getTau(self, data, l):
        return self.quotient(data, l, 1 / (2 * self.Const))


class MultiPassiveAggressiveFilter(PassiveAggressiveFilter):
    """The reference's multiclass legacy filter is an abstract stub
    (algorithm/multi/PassiveAggressiveFilter.scala); use ``models.pa.algorithms`` for multiclass."""


class _TrainEOF(EOF):
    pass


class _TestEOF(EOF):
    pass


class _OfflineWorker(CtorFork, WorkerLogic):
    def __init__(self, paFilter: PassiveAggressiveFilter, iterations: int, shuffle: bool = False,
                 seed: Optional[int] = None):
        self.paf, self.iterations, self.shuffle = paFilter, iterations, shuffle
        self.train: List[Tuple[LegacySparseVector, int]] = []
        self.test: List[LegacySparseVector] = []
        self.waiting: Dict[int, deque] = {}
        self.lock = threading.Lock()
        self.thread: Optional[threading.Thread] = None
        self.train_done = self.test_done = False
        self.rnd = random.Random(seed)

    def _request(self, vec, label, ps):
        pending = {"vec": vec, "label": label, "values": {}}
        with self.lock:
            for k in vec.getIndexes():
                self.waiting.setdefault(k, deque()).append(pending)
        for k in list(vec.getIndexes()):
            ps.pull(k)

    def _start(self, ps):
        def run():
            for _ in range(self.iterations):
                if self.shuffle:
                    self.rnd.shuffle(self.train)
                for vec, label in self.train:
                    self._request(vec, label, ps)
            for vec in self.test:
                self._request(vec, None, ps)

        self.thread = threading.Thread(target=run, daemon=True, name="fps-pa-offline")
        self.thread.start()

    def onRecv(self, data, ps):
        if isinstance(data, _TrainEOF):
            self.train_done = True
        elif isinstance(data, _TestEOF):
            self.test_done = True
        elif data[1] is None:
            self.test.append(data[0])
        else:
            self.train.append((data[0], int(data[1])))
        if self.train_done and self.test_done and self.thread is None:
            self._start(ps)

    def onPullRecv(self, paramId, value, ps):
        with self.lock:
            q = self.waiting[paramId]
            pending = q.popleft()
            if not q:
                del self.waiting[paramId]
        pending["values"][paramId] = value
        vec = pending["vec"]
        if len(pending["values"]) == len(vec.getIndexes()):
            if pending["label"] is None:
                ps.output((vec, self.paf.predict(vec, pending["values"])))
            else:
                for k, d in self.paf.delta(vec, pending["values"], pending["label"]).items():
                    ps.push(k, d)

    def close(self):
        if self.thread is not None:
            self.thread.join(timeout=60)


def paBinaryClassificationOffline(trainingSrc, testSrc, workerParallelism: int, psParallelism: int,
                                  iterations: int, pafType: int = 0, pafConst: float = 1,
                                  pullLimit: int = 10000, iterationWaitTime: float = 10000,
                                  seed: Optional[int] = None, shuffle: bool = False):
    """``trainingSrc``: ``(LegacySparseVector, +1/-1)`` records; ``testSrc``: ``LegacySparseVector`` or
    ``(vector, _)`` records.  ``pafType``: 0 = PA, 1 = PA-I(pafConst), 2 = PA-II(pafConst)."""
    paf = {0: PassiveAggressiveFilter.buildPAF, 1: lambda: PassiveAggressiveFilter.buildPAFI(pafConst),
           2: lambda: PassiveAggressiveFilter.buildPAFII(pafConst)}[pafType]()
    rnd = random.Random(seed)
    shuffle_part = lambda _k, n: rnd.randrange(n)
    train = as_stream(trainingSrc).map(lambda r: (r[0], r[1])).partition_custom(shuffle_part).with_eof(_TrainEOF)
    test = as_stream(testSrc).map(lambda r: (r[0] if isinstance(r, tuple) else r, None)) \
        .partition_custom(shuffle_part).with_eof(_TestEOF)
    worker = addBlockingPullLimiter(_OfflineWorker(paf, iterations, shuffle, seed), pullLimit)
    return transform(train.union(test), worker, SimplePSLogicWithClose(lambda _i: 0.0, lambda a, b: a + b),
                     workerParallelism, psParallelism, iterationWaitTime)


paMultiClassificationOffline = paBinaryClassificationOffline  # the reference's copy differs by name only
