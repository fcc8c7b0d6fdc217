"""Passive-aggressive classifiers on the parameter server -- ``transformBinary``,
``transformMulticlass``, ``transformMulticlassWithLongId``.

Reference: M/passive/aggressive/PassiveAggressiveParameterServer.scala:63-461.  One PS parameter
per feature id (binary: a float; multiclass: a ``labelCount`` vector).  ``onRecv`` pulls every
active feature of the example; when all answers arrived the worker builds the model slice and
either pushes the per-feature deltas (labelled, ``Left((vector, label))``) or outputs the
prediction (unlabelled, ``Right((id, vector))``).  Optional initial model (-> model load), hash or
range partitioning, always behind a pull limiter; the model is dumped at close.

``backend="device"`` runs the same algorithm as one fused CSR kernel per micro-batch
(models/pa/device.py, ops/csrc/fps_pa.cu).
"""
from __future__ import annotations

import math
from collections import deque
from typing import Any, Dict, Optional

from ...api import Left, Right, WorkerLogic
from ...limiter import addPullLimiter
from ...runtime.transform import (default_worker_partitioner, transform_general,
                                  transformWithModelLoad)
from ...protocol.senders import (SimplePSReceiver, SimplePSSender, SimpleWorkerReceiver,
                                 SimpleWorkerSender)
from ...server.logics import RangePSLogicWithClose, SimplePSLogicWithClose
from .algorithms import PassiveAggressiveAlgorithm, initBinary, initMulti
from .sparse import SparseVector


def rangePartitionerPS(featureCount: int, psParallelism: int):
    """``abs(id) // ceil(featureCount / psParallelism)`` (PassiveAggressiveParameterServer.scala:374-386)."""
    size = int(math.ceil(featureCount / psParallelism))

    def part(msg) -> int:
        m = msg[0] if isinstance(msg, (list, tuple)) else msg
        return abs(int(m.paramId)) // size

    return part


def hashPartitionerPS(psParallelism: int):
    def part(msg) -> int:
        m = msg[0] if isinstance(msg, (list, tuple)) else msg
        return abs(int(m.paramId)) % psParallelism

    return part


class _Pending:
    __slots__ = ("data", "values")

    def __init__(self, data):
        self.data = data
        self.values: Dict[int, Any] = {}


class PassiveAggressiveWorkerLogic(WorkerLogic):
    """Gather-all-features-then-act worker (PassiveAggressiveParameterServer.scala:283-340)."""

    def __init__(self, algo: PassiveAggressiveAlgorithm, id_of):
        self.algo = algo
        self.id_of = id_of
        self.paramWaitingQueue: Dict[int, deque] = {}

    @staticmethod
    def _vector(data) -> SparseVector:
        return data.value[0] if data.is_left else data.value[1]

    def onRecv(self, data, ps):
        vec = self._vector(data)
        pending = _Pending(data)
        for k in vec.indices.tolist():
            self.paramWaitingQueue.setdefault(k, deque()).append(pending)
            ps.pull(k)

    def onPullRecv(self, paramId, modelValue, ps):
        q = self.paramWaitingQueue[paramId]
        pending = q.popleft()
        pending.values[paramId] = modelValue
        vec = self._vector(pending.data)
        if len(pending.values) == vec.activeSize:
            if pending.data.is_left:
                label = pending.data.value[1]
                for i, v in self.algo.delta(vec, pending.values, label):
                    ps.push(i, v)
            else:
                ps.output((self.id_of(pending.data), self.algo.predict(vec, pending.values)))
        if not q:
            del self.paramWaitingQueue[paramId]


def _vec_as_id(data):
    return data.value[1]


def _long_id(data):
    if data.is_left:
        raise ValueError("Only unlabelled vectors have id.")
    return data.value[0]


def _transform_generic(model, init, add, inputSource, workerParallelism, psParallelism, algo,
                       pullLimit, featureCount, rangePartitioning, iterationWaitTime, id_of):
    serverLogic = (RangePSLogicWithClose(featureCount, init, add) if rangePartitioning
                   else SimplePSLogicWithClose(init, add))
    paramPartitioner = (rangePartitionerPS(featureCount, psParallelism) if rangePartitioning
                        else hashPartitionerPS(psParallelism))
    workerLogic = addPullLimiter(PassiveAggressiveWorkerLogic(algo, id_of), pullLimit)
    wIn = default_worker_partitioner(workerParallelism)
    if model is not None:
        return transformWithModelLoad(model)(inputSource, workerLogic, serverLogic, paramPartitioner,
                                             wIn, workerParallelism, psParallelism, iterationWaitTime)
    return transform_general(inputSource, workerLogic, serverLogic, paramPartitioner, wIn,
                             workerParallelism, psParallelism, SimpleWorkerReceiver(),
                             SimpleWorkerSender(), SimplePSReceiver(), SimplePSSender(),
                             iterationWaitTime)


def _transform_binary_native(model, inputSource, workerParallelism, psParallelism, algo, pullLimit, featureCount,
                             rangePartitioning):
    """``backend="native"``: the same job on the C++ host engine (worker / server threads over SPSC rings,
    ``ops/csrc/fps_host.cpp::fps_host_pa_binary``)."""
    import numpy as np

    from ...ops import host
    from ...runtime.stream import ResultStream, as_stream
    from .device import algo_to_device

    name, C, _ = algo_to_device(algo)
    recs = as_stream(inputSource).collect()
    weights = np.zeros(int(featureCount), dtype=np.float32)
    if model is not None:
        for fid, w in as_stream(model).collect():
            weights[int(fid)] = w
    row_ptr, cols, vals, labels = [0], [], [], []
    for d in recs:
        vec = d.value[0] if d.is_left else d.value[1]
        cols.extend(vec.indices.tolist()); vals.extend(vec.values.tolist())
        row_ptr.append(len(cols))
        labels.append((1 if d.value[1] else -1) if d.is_left else 0)
    pred, weights, touched = host.pa_binary(row_ptr, cols, vals, labels, int(featureCount), name, C,
                                            workers=workerParallelism, servers=psParallelism,
                                            pull_limit=max(1, int(pullLimit)),
                                            range_partitioning=bool(rangePartitioning), weights=weights)
    out = [Left((d.value[1], bool(p))) for d, p in zip(recs, pred) if not d.is_left]
    keep = touched | (weights != 0)
    out += [Right((int(f), float(weights[f]))) for f in np.nonzero(keep)[0]]
    return ResultStream(out)


def _transform_multiclass_native(model, inputSource, workerParallelism, psParallelism, algo, pullLimit, labelCount,
                                 featureCount, rangePartitioning):
    """``backend="native"`` of :func:`transformMulticlass` (``fps_host_pa_multiclass``)."""
    import numpy as np

    from ...ops import host
    from ...runtime.stream import ResultStream, as_stream
    from .device import algo_to_device

    name, C, cost = algo_to_device(algo)
    L = int(labelCount)
    recs = as_stream(inputSource).collect()
    weights = np.zeros((int(featureCount), L), dtype=np.float32)
    if model is not None:
        for fid, w in as_stream(model).collect():
            weights[int(fid)] = np.asarray(w, dtype=np.float32)
    row_ptr, cols, vals, labels = [0], [], [], []
    for d in recs:
        vec = d.value[0] if d.is_left else d.value[1]
        cols.extend(vec.indices.tolist()); vals.extend(vec.values.tolist())
        row_ptr.append(len(cols))
        labels.append(int(d.value[1]) if d.is_left else -1)
    pred, weights, touched = host.pa_multiclass(row_ptr, cols, vals, labels, int(featureCount), L, name, C, cost,
                                                workers=workerParallelism, servers=psParallelism,
                                                pull_limit=max(1, int(pullLimit)),
                                                range_partitioning=bool(rangePartitioning), weights=weights)
    out = [Left((d.value[1], int(p))) for d, p in zip(recs, pred) if not d.is_left]
    keep = touched | (weights != 0).any(axis=1)
    out += [Right((int(f), weights[f].astype("float64"))) for f in np.nonzero(keep)[0]]
    return ResultStream(out)


def transformBinary(model=None):
    """``transformBinary(model)(inputSource, workerParallelism, psParallelism, algo, pullLimit,
    featureCount, rangePartitioning, iterationWaitTime)``; input records are
    ``Left((SparseVector, bool))`` (train) or ``Right((id, SparseVector))`` (predict); output
    ``Left((vector, predictedLabel))`` / ``Right((featureId, weight))``."""

    def run(inputSource, workerParallelism, psParallelism, passiveAggressiveMethod, pullLimit,
            featureCount, rangePartitioning=False, iterationWaitTime=10000, backend="local", **kw):
        if backend == "native":
            return _transform_binary_native(model, inputSource, workerParallelism, psParallelism,
                                            passiveAggressiveMethod, pullLimit, featureCount, rangePartitioning)
        if backend == "device":
            from .device import transform_binary_device

            return transform_binary_device(model, inputSource, passiveAggressiveMethod,
                                           featureCount, rangePartitioning, pullLimit=pullLimit, **kw)
        return _transform_generic(model, initBinary, lambda a, b: a + b, inputSource,
                                  workerParallelism, psParallelism, passiveAggressiveMethod, pullLimit,
                                  featureCount, rangePartitioning, iterationWaitTime, _vec_as_id)

    return run


def transformMulticlass(model=None):
    """Multiclass variant: parameters are ``labelCount`` vectors; vector itself is the output id."""

    def run(inputSource, workerParallelism, psParallelism, passiveAggressiveMethod, pullLimit,
            labelCount, featureCount, rangePartitioning=False, iterationWaitTime=10000,
            backend="local", **kw):
        if backend == "native":
            return _transform_multiclass_native(model, inputSource, workerParallelism, psParallelism,
                                                passiveAggressiveMethod, pullLimit, labelCount, featureCount,
                                                rangePartitioning)
        if backend == "device":
            from .device import transform_multiclass_device

            return transform_multiclass_device(model, inputSource, passiveAggressiveMethod, labelCount,
                                               featureCount, rangePartitioning, pullLimit=pullLimit, **kw)
        return _transform_generic(model, initMulti(labelCount), lambda a, b: a + b, inputSource,
                                  workerParallelism, psParallelism, passiveAggressiveMethod, pullLimit,
                                  featureCount, rangePartitioning, iterationWaitTime, _vec_as_id)

    return run


def transformMulticlassWithLongId(model=None):
    """Like :func:`transformMulticlass` but unlabelled records carry a long id that is echoed in
    the prediction output."""

    def run(inputSource, workerParallelism, psParallelism, passiveAggressiveMethod, pullLimit,
            labelCount, featureCount, rangePartitioning=False, iterationWaitTime=10000):
        return _transform_generic(model, initMulti(labelCount), lambda a, b: a + b, inputSource,
                                  workerParallelism, psParallelism, passiveAggressiveMethod, pullLimit,
                                  featureCount, rangePartitioning, iterationWaitTime, _long_id)

    return run


transform_binary = transformBinary
transform_multiclass = transformMulticlass
transform_multiclass_with_long_id = transformMulticlassWithLongId
