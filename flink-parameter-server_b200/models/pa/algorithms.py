"""Passive-aggressive update rules (host tier; device equivalents in ops/csrc/fps_pa.cu).

* binary PA / PA-I / PA-II  (M/passive/aggressive/algorithm/PassiveAggressiveBinaryAlgorithm.scala:11-112)
* multiclass one-versus-all PA / PA-I / PA-II  (PassiveAggressiveOneVersusAll.scala:14-123)
* cost-sensitive multiclass PB / ML  (PassiveAggressiveCostBased.scala:13-140); the reference
  reuses one ``VectorBuilder`` across features without reset (quirk) -- here every
  feature gets its own fresh delta.

A *model* is what the worker assembled from the pulled parameters: for binary a
``{featureId: weight}`` dict, for multiclass a ``{featureId: ndarray[labelCount]}`` dict.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Tuple

import numpy as np

from .sparse import SparseVector


class PassiveAggressiveAlgorithm:
    """``delta(dataPoint, model, label) -> [(featureId, paramDelta)]``, ``predict(dataPoint, model)``
    (PassiveAggressiveAlgorithm.scala)."""

    def delta(self, dataPoint: SparseVector, model, label):
        raise NotImplementedError

    def predict(self, dataPoint: SparseVector, model):
        raise NotImplementedError


def initBinary(_featureId: int) -> float:
    return 0.0


def initMulti(labelCount: int) -> Callable[[int], np.ndarray]:
    return lambda _i: np.zeros(labelCount)


class RandomModelInitializer:
    """Despite its name the reference returns 0 (RandomModelInitializer.scala)."""

    @staticmethod
    def init() -> float:
        return 0.0


# ---- binary -------------------------------------------------------------------------------
class PassiveAggressiveBinaryAlgorithm(PassiveAggressiveAlgorithm):
    def __init__(self, aggressiveness: float = 0.0):
        self.aggressiveness = aggressiveness

    @staticmethod
    def buildPA():
        return PassiveAggressiveBinaryAlgorithmImpl()

    @staticmethod
    def buildPAI(aggressiveness: float):
        return PassiveAggressiveBinaryAlgorithmImplI(aggressiveness)

    @staticmethod
    def buildPAII(aggressiveness: float):
        return PassiveAggressiveBinaryAlgorithmImplII(aggressiveness)

    def tau(self, dataPoint: SparseVector, loss: float) -> float:
        raise NotImplementedError

    @staticmethod
    def quotient(dataPoint: SparseVector, loss: float, denominatorConst: float) -> float:
        n2 = dataPoint.norm_sq()
        return loss / n2 if denominatorConst == 0 else loss / (n2 + denominatorConst)

    def delta(self, dataPoint, model: Dict[int, float], label: bool) -> List[Tuple[int, float]]:
        y = 1 if label else -1
        loss = max(0.0, 1 - y * dataPoint.dot(model))
        mult = self.tau(dataPoint, loss) * y
        if mult == 0.0:
            return []
        return [(i, v * mult) for i, v in dataPoint.activeIterator()]

    def predict(self, dataPoint, model) -> bool:
        return dataPoint.dot(model) > 0


class PassiveAggressiveBinaryAlgorithmImpl(PassiveAggressiveBinaryAlgorithm):
    def tau(self, d, loss):
        return self.quotient(d, loss, 0)


class PassiveAggressiveBinaryAlgorithmImplI(PassiveAggressiveBinaryAlgorithm):
    def tau(self, d, loss):
        return min(self.aggressiveness, self.quotient(d, loss, 0))


class PassiveAggressiveBinaryAlgorithmImplII(PassiveAggressiveBinaryAlgorithm):
    def tau(self, d, loss):
        return self.quotient(d, loss, 1 / (2 * self.aggressiveness))


# ---- multiclass -----------------------------------------------------------------------------
def _decision(dataPoint: SparseVector, model: Dict[int, np.ndarray], labelCount: int) -> np.ndarray:
    d = np.zeros(labelCount)
    for i, v in dataPoint.activeIterator():
        w = model.get(i)
        if w is not None:
            d += v * w
    return d


class PassiveAggressiveMulticlassAlgorithm(PassiveAggressiveAlgorithm):
    labelCount: int

    def predict(self, dataPoint, model) -> int:
        return int(np.argmax(_decision(dataPoint, model, self.labelCount)))


class PassiveAggressiveOneVersusAll(PassiveAggressiveMulticlassAlgorithm):
    def __init__(self, labelCount: int, aggressiveness: float = 0.0):
        self.labelCount = labelCount
        self.aggressiveness = aggressiveness

    @staticmethod
    def buildPA(labelCount: int):
        return PassiveAggressiveOneVersusAllImpl(labelCount)

    @staticmethod
    def buildPAI(labelCount: int, aggressiveness: float):
        return PassiveAggressiveOneVersusAllImplI(labelCount, aggressiveness)

    @staticmethod
    def buildPAII(labelCount: int, aggressiveness: float):
        return PassiveAggressiveOneVersusAllImplII(labelCount, aggressiveness)

    def tau(self, normSquare: float, loss: np.ndarray) -> np.ndarray:
        raise NotImplementedError

    @staticmethod
    def loss(decisionVector: np.ndarray, labelVect: np.ndarray) -> np.ndarray:
        return np.maximum(0.0, 1 - decisionVector * labelVect)

    def delta(self, dataPoint, model, label: int):
        labelVector = -np.ones(self.labelCount)
        labelVector[label] = 1
        mult = self.tau(dataPoint.norm_sq(),
                        self.loss(_decision(dataPoint, model, self.labelCount), labelVector)) * labelVector
        if not mult.any():
            return []
        return [(i, v * mult) for i, v in dataPoint.activeIterator()]


class PassiveAggressiveOneVersusAllImpl(PassiveAggressiveOneVersusAll):
    def tau(self, n2, loss):
        return loss / n2


class PassiveAggressiveOneVersusAllImplI(PassiveAggressiveOneVersusAll):
    def tau(self, n2, loss):
        return np.minimum(self.aggressiveness, loss / n2)


class PassiveAggressiveOneVersusAllImplII(PassiveAggressiveOneVersusAll):
    def tau(self, n2, loss):
        return loss / (n2 + 1 / (2 * self.aggressiveness))


class PassiveAggressiveCostBased(PassiveAggressiveMulticlassAlgorithm):
    def __init__(self, cost: Callable[[int, int], float], labelCount: int):
        self.cost = cost
        self.labelCount = labelCount

    @staticmethod
    def buildPB(cost, labelCount):
        return PassiveAggressiveCostBasedImplPB(cost, labelCount)

    @staticmethod
    def buildML(cost, labelCount):
        return PassiveAggressiveCostBasedImplML(cost, labelCount)

    def quotient(self, decisionVector: np.ndarray, label: int) -> int:
        raise NotImplementedError

    def loss(self, d: np.ndarray, q: int, label: int) -> float:
        return float(d[q] - d[label] + math.sqrt(self.cost(label, q)))

    @staticmethod
    def tau(dataPoint: SparseVector, loss: float) -> float:
        return loss / (2 * dataPoint.norm_sq())

    def delta(self, dataPoint, model, label: int):
        d = _decision(dataPoint, model, self.labelCount)
        q = self.quotient(d, label)
        if q == label:
            return []
        t = self.tau(dataPoint, self.loss(d, q, label))
        out = []
        for i, v in dataPoint.activeIterator():
            dv = np.zeros(self.labelCount)
            dv[label] += t * v
            dv[q] -= t * v
            out.append((i, dv))
        return out


class PassiveAggressiveCostBasedImplPB(PassiveAggressiveCostBased):
    """Prediction-based: ``q = argmax d``."""

    def quotient(self, d, label):
        return int(np.argmax(d))


class PassiveAggressiveCostBasedImplML(PassiveAggressiveCostBased):
    """Max-loss: ``q = argmax_i d_i - d_label + sqrt(cost(label, i))``."""

    def quotient(self, d, label):
        s = np.array([d[i] - d[label] + math.sqrt(self.cost(label, i)) for i in range(self.labelCount)])
        return int(np.argmax(s))
