"""Shared matrix-factorisation building blocks (host tier).

Capabilities of M/matrix/factorization/{factors,utils}/: vector helpers, ``Rating`` / ``RichRating``,
factor initialisers (incl. the descriptor/``open()`` factory pattern), ``SGDUpdater``, id generator,
top-K queue and the partitioner pair.  Vectors are ``numpy.float64`` arrays (the reference's
``Array[Double]``); the device tier uses fp32 rows in HBM.
"""
from __future__ import annotations

import heapq
import itertools
import math
import random
import threading
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from ...parallel.partitioner import stable_hash
from ...utils.input_source import EventWithTimestamp

Vector = np.ndarray
UserId = int
ItemId = int


from ...errors import FactorIsNotANumberException  # noqa: E402,F401  (re-exported; Vector.scala:78-80)


def vectorLengthSqr(v: Vector) -> float:
    return float(np.dot(v, v))


def dotProduct(u: Vector, v: Vector) -> float:
    return float(np.dot(u, v))


def vectorSum(u: Vector, v: Vector) -> Vector:
    res = u + v
    if np.isnan(res).any():
        raise FactorIsNotANumberException()
    return res


def attachLength(u: Vector) -> Tuple[float, Vector]:
    return math.sqrt(vectorLengthSqr(u)), u


@dataclass(frozen=True)
class Rating(EventWithTimestamp):
    user: int
    item: int
    rating: float
    timestamp: int = 0

    def enrich(self, workerId: int, ratingId: int) -> "RichRating":
        return RichRating(self, workerId, ratingId)

    def getEventTime(self) -> int:
        return self.timestamp


@dataclass(frozen=True)
class RichRating(EventWithTimestamp):
    """A rating with a target worker and a rating id (InputTypes.scala:44-49)."""

    base: Rating
    targetWorker: int
    ratingId: int

    def reduce(self) -> Rating:
        return self.base

    def getEventTime(self) -> int:
        return self.base.getEventTime()


def ratingFromTuple(t) -> Rating:
    return Rating(int(t[0]), int(t[1]), float(t[2]), 0)


class IDGenerator:
    _n = itertools.count()
    _lock = threading.Lock()

    @classmethod
    def next(cls) -> int:
        with cls._lock:
            return next(cls._n)


# ---- factor initialisers ------------------------------------------------------------------
class FactorInitializer:
    def nextFactor(self, id: int) -> Vector:
        raise NotImplementedError


class FactorInitializerDescriptor:
    """Serializable recipe; ``open()`` builds the (non-serialisable RNG holding) initialiser."""

    def open(self) -> FactorInitializer:
        raise NotImplementedError


class RandomFactorInitializer(FactorInitializer):
    """U[0,1) (RandomFactorInitializer.scala)."""

    def __init__(self, rnd: random.Random, numFactors: int):
        self.rnd, self.n = rnd, numFactors

    def nextFactor(self, id):
        return np.array([self.rnd.random() for _ in range(self.n)])


class RandomFactorInitializerDescriptor(FactorInitializerDescriptor):
    def __init__(self, numFactors: int):
        self.numFactors = numFactors

    def open(self):
        return RandomFactorInitializer(random.Random(), self.numFactors)


class RangedRandomFactorInitializer(FactorInitializer):
    """U[min,max) (RangedRandomFactorInitializer.scala:7-9)."""

    def __init__(self, rnd: random.Random, numFactors: int, rangeMin: float, rangeMax: float):
        self.rnd, self.n, self.lo, self.hi = rnd, numFactors, rangeMin, rangeMax

    def nextFactor(self, id):
        return np.array([self.lo + (self.hi - self.lo) * self.rnd.random() for _ in range(self.n)])


class RangedRandomFactorInitializerDescriptor(FactorInitializerDescriptor):
    def __init__(self, numFactors: int, rangeMin: float, rangeMax: float, seed: Optional[int] = None):
        self.numFactors, self.rangeMin, self.rangeMax, self.seed = numFactors, rangeMin, rangeMax, seed

    def open(self):
        return RangedRandomFactorInitializer(random.Random(self.seed), self.numFactors,
                                             self.rangeMin, self.rangeMax)


class PseudoRandomFactorInitializer(FactorInitializer):
    """Deterministic: RNG seeded by the id (PseudoRandomFactorInitializer.scala:7-20)."""

    def __init__(self, numFactors: int):
        self.n = numFactors

    def nextFactor(self, id):
        r = random.Random(int(id))
        return np.array([r.random() for _ in range(self.n)])


class PseudoRandomFactorInitializerDescriptor(FactorInitializerDescriptor):
    def __init__(self, numFactors: int):
        self.numFactors = numFactors

    def open(self):
        return PseudoRandomFactorInitializer(self.numFactors)


# ---- updaters ---------------------------------------------------------------------------
class FactorUpdater:
    def delta(self, rating: float, user: Vector, item: Vector) -> Tuple[Vector, Vector]:
        raise NotImplementedError


def _sigmoid(x: float) -> float:
    if x >= 0:
        return 1.0 / (1.0 + math.exp(-x))
    e = math.exp(x)
    return e / (1.0 + e)


class SGDUpdater(FactorUpdater):
    """``e = sigmoid(r - u.v)`` (parity, SGDUpdater.scala:8) or plain residual; no regulariser."""

    def __init__(self, learningRate: float, plain_residual: bool = False):
        self.lr = learningRate
        self.plain = plain_residual

    def delta(self, rating, user, item):
        resid = rating - float(np.dot(user, item))
        e = resid if self.plain else _sigmoid(resid)
        return self.lr * e * item, self.lr * e * user


# ---- top-K ------------------------------------------------------------------------------
class TopKQueue:
    """Bounded min-heap of ``(score, itemId)`` keeping the K largest (Utils.scala:13-18)."""

    def __init__(self, k: Optional[int] = None):
        self.k = k
        self.h: List[Tuple[float, int]] = []

    def push(self, score: float, item: int) -> None:
        if self.k is None or len(self.h) < self.k:
            heapq.heappush(self.h, (score, item))
        elif (score, item) > self.h[0]:
            heapq.heapreplace(self.h, (score, item))

    def min_score(self) -> float:
        return self.h[0][0] if self.h else -math.inf

    def __len__(self):
        return len(self.h)

    def sorted_desc(self) -> List[Tuple[float, int]]:
        return sorted(self.h, reverse=True)


class Partitioner:
    """``hash(id) % psParallelism`` / answer-to-worker pair (Utils.scala:44-63)."""

    def __init__(self, psParallelism: int):
        self.psP = psParallelism

    def workerToPSPartitioner(self, msg) -> int:
        m = msg[0] if isinstance(msg, (list, tuple)) else msg
        return stable_hash(m.paramId) % self.psP

    @staticmethod
    def psToWorkerPartitioner(msg) -> int:
        m = msg[0] if isinstance(msg, (list, tuple)) else msg
        return m.workerPartitionIndex
