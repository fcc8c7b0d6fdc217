"""Sparse vectors for the passive-aggressive classifiers.

``SparseVector`` plays the role of Breeze's ``SparseVector[Double]`` (used by
M/passive/aggressive/PassiveAggressiveParameterServer.scala) and of the reference's own legacy
``entities/SparseVector.scala:6-42`` (``LegacySparseVector`` keeps that map-based API).
"""
from __future__ import annotations

from typing import Dict, Iterable, Iterator, Mapping, Tuple

import numpy as np


class SparseVector:
    __slots__ = ("indices", "values", "length")

    def __init__(self, indices, values, length: int):
        idx = np.asarray(indices, dtype=np.int64)
        val = np.asarray(values, dtype=np.float64)
        order = np.argsort(idx, kind="stable")
        self.indices, self.values, self.length = idx[order], val[order], int(length)

    @staticmethod
    def from_dict(d: Mapping[int, float], length: int) -> "SparseVector":
        return SparseVector(list(d.keys()), list(d.values()), length)

    @property
    def activeSize(self) -> int:
        return int(self.indices.size)

    def activeIterator(self) -> Iterator[Tuple[int, float]]:
        return zip(self.indices.tolist(), self.values.tolist())

    def dot(self, other) -> float:
        if isinstance(other, SparseVector):
            i, ia, ib = np.intersect1d(self.indices, other.indices, return_indices=True)
            return float(np.dot(self.values[ia], other.values[ib]))
        if isinstance(other, dict):
            return float(sum(v * other.get(int(i), 0.0) for i, v in self.activeIterator()))
        return float(np.dot(self.values, np.asarray(other)[self.indices]))

    def norm_sq(self) -> float:
        return float(np.dot(self.values, self.values))

    def __mul__(self, s: float) -> "SparseVector":
        return SparseVector(self.indices, self.values * s, self.length)

    def __eq__(self, o) -> bool:
        return (isinstance(o, SparseVector) and self.length == o.length
                and np.array_equal(self.indices, o.indices) and np.array_equal(self.values, o.values))

    def __hash__(self) -> int:
        return hash((self.length, self.indices.tobytes(), self.values.tobytes()))

    def __repr__(self) -> str:
        return f"SparseVector(n={self.length}, nnz={self.activeSize})"


class LegacySparseVector:
    """HashMap-backed sparse vector of the legacy offline PA pipeline
    (M/passive/aggressive/entities/SparseVector.scala:6-42)."""

    def __init__(self, indexes: Dict[int, float], default: float = 0.0):
        self.indexes = dict(indexes)
        self.default = default

    @staticmethod
    def build(pairs: Iterable[Tuple[int, float]], default: float = 0.0) -> "LegacySparseVector":
        return LegacySparseVector(dict(pairs), default)

    @staticmethod
    def endOfFile(workerId: int, minusSourceId: int, default: float = 0.0) -> "EOFSign":
        return EOFSign(workerId, minusSourceId)

    def getIndexes(self):
        return self.indexes.keys()

    def getValues(self):
        return dict(self.indexes)

    def get(self, i: int) -> float:
        return self.indexes.get(i, self.default)

    def __eq__(self, o) -> bool:
        return isinstance(o, LegacySparseVector) and self.indexes == o.indexes

    def __hash__(self) -> int:
        return hash(frozenset(self.indexes.items()))


class EOFSign(LegacySparseVector):
    """End-of-input marker of the legacy offline pipeline: an empty vector that carries the worker it is
    addressed to and the (negated) id of the source that finished (entities/SparseVector.scala:13,42)."""

    def __init__(self, workerId: int, minusSourceId: int):
        super().__init__({}, 0.0)
        self.workerId, self.minusSourceId = int(workerId), int(minusSourceId)

    def __eq__(self, o) -> bool:
        return isinstance(o, EOFSign) and (self.workerId, self.minusSourceId) == (o.workerId, o.minusSourceId)

    def __hash__(self) -> int:
        return hash(("EOFSign", self.workerId, self.minusSourceId))

    def __repr__(self) -> str:
        return f"EOFSign(workerId={self.workerId}, minusSourceId={self.minusSourceId})"
