"""Passive-aggressive classifiers (host tier): algorithm unit checks + the reference's quality gate
(T/passive/aggressive/PassiveAggressiveParameterServerTest.scala:44-100: accuracy >= 80 %, scaled down)."""
import random

import pytest

import numpy as np

from fps_b200.api import Left, Right
from fps_b200.models.pa.algorithms import (PassiveAggressiveBinaryAlgorithm, PassiveAggressiveCostBased,
                                           PassiveAggressiveOneVersusAll)
from fps_b200.models.pa.ps import transformBinary, transformMulticlass, transformMulticlassWithLongId
from fps_b200.models.pa.sparse import EOFSign, LegacySparseVector, SparseVector


def test_sparse_vector_and_legacy_builders():
    v = SparseVector([5, 1], [2.0, 3.0], 10)
    assert v.indices.tolist() == [1, 5] and v.dot({1: 1.0, 5: 0.5}) == 4.0 and v.norm_sq() == 13.0
    a = LegacySparseVector.build([(1, 2.0), (3, 4.0)])
    assert a == LegacySparseVector({1: 2.0, 3: 4.0}) and a.get(9) == 0.0
    eof = LegacySparseVector.endOfFile(2, -7)              # entities/SparseVector.scala:13,42
    assert isinstance(eof, EOFSign) and (eof.workerId, eof.minusSourceId) == (2, -7) and eof.getValues() == {}
    assert eof == EOFSign(2, -7) and eof != EOFSign(2, -8)


def test_binary_pa_variants():
    x = SparseVector([0, 2], [1.0, 2.0], 4)       # ||x||^2 = 5
    w = {0: 0.0, 2: 0.0}
    assert PassiveAggressiveBinaryAlgorithm.buildPA().delta(x, w, True) == [(0, 0.2), (2, 0.4)]
    assert PassiveAggressiveBinaryAlgorithm.buildPAI(0.1).delta(x, w, False) == [(0, -0.1), (2, -0.2)]
    d = PassiveAggressiveBinaryAlgorithm.buildPAII(1.0).delta(x, w, True)
    np.testing.assert_allclose([v for _, v in d], [1 / 5.5, 2 / 5.5])
    assert PassiveAggressiveBinaryAlgorithm.buildPA().delta(x, {0: 10.0, 2: 0.0}, True) == []  # no loss
    assert PassiveAggressiveBinaryAlgorithm.buildPA().predict(x, {0: 1.0, 2: -0.1}) is True


def test_multiclass_ova_and_cost_based():
    x = SparseVector([1], [2.0], 3)
    m = {1: np.zeros(3)}
    d = PassiveAggressiveOneVersusAll.buildPA(3).delta(x, m, 2)
    np.testing.assert_allclose(d[0][1], 2.0 * np.array([-1, -1, 1]) / 4.0)
    cost = lambda a, b: 0.0 if a == b else 1.0
    pb = PassiveAggressiveCostBased.buildPB(cost, 3)
    m2 = {1: np.array([1.0, 0.0, 0.0])}
    d2 = pb.delta(x, m2, 2)                     # q = argmax = 0, loss = 2 - 0 + 1 = 3, tau = 3/8
    np.testing.assert_allclose(d2[0][1], np.array([-0.75, 0.0, 0.75]))
    assert pb.delta(x, {1: np.array([0.0, 0.0, 5.0])}, 2) == []
    ml = PassiveAggressiveCostBased.buildML(cost, 3)
    assert ml.quotient(np.array([0.0, 0.0, 5.0]), 2) == 2


def _dataset(n, feats, nnz, seed):
    r = random.Random(seed)
    w_true = np.array([r.gauss(0, 1) for _ in range(feats)])
    out = []
    for _ in range(n):
        idx = r.sample(range(feats), nnz)
        val = [r.gauss(0, 1) for _ in idx]
        v = SparseVector(idx, val, feats)
        out.append((v, v.dot(w_true) > 0))
    return out


def test_binary_pa_accuracy_gate_range_partitioning_3x3():
    feats, data = 2000, _dataset(80, 2000, 200, seed=7)
    train = [Left(d) for d in data] * 3
    out = transformBinary()(train, 3, 3, PassiveAggressiveBinaryAlgorithm.buildPA(), 500, feats, True, 200)
    model = dict(out.ps_outputs())
    acc = sum((v.dot(model) > 0) == y for v, y in data[:20]) / 20
    assert acc >= 0.8, acc
    # the reference's evaluation helper (percent; refuses unlabelled examples)
    from fps_b200.models.pa.evaluation import PassiveAggressiveBinaryModelEvaluation as Ev
    pac = PassiveAggressiveBinaryAlgorithm.buildPA()
    assert abs(Ev.accuracy(model, data[:20], feats, pac) - 100.0 * acc) < 1e-9
    c = Ev.confusion(model, data[:20], pac)
    assert sum(c.values()) == 20 and c["tt"] + c["ff"] == round(acc * 20)
    with pytest.raises(ValueError):
        Ev.accuracy(model, [(data[0][0], None)], feats, pac)
    # predict path with the trained model loaded back through transformWithModelLoad
    pred = transformBinary(list(model.items()))([Right((i, v)) for i, (v, _) in enumerate(data[:20])],
                                                3, 3, PassiveAggressiveBinaryAlgorithm.buildPA(), 500,
                                                feats, True, 200)
    got = {v: p for v, p in pred.worker_outputs()}
    acc2 = sum(got[v] == y for v, y in data[:20]) / 20
    assert acc2 >= 0.8


def test_multiclass_pa_with_long_ids():
    r = random.Random(3)
    feats, L = 300, 3
    protos = [np.array([r.gauss(0, 1) for _ in range(feats)]) for _ in range(L)]
    data = []
    for _ in range(150):
        idx = r.sample(range(feats), 40)
        v = SparseVector(idx, [r.gauss(0, 1) for _ in idx], feats)
        data.append((v, int(np.argmax([v.dot(p) for p in protos]))))
    algo = PassiveAggressiveOneVersusAll.buildPAI(L, 1.0)
    out = transformMulticlass()([Left(d) for d in data] * 2, 2, 2, algo, 100, L, feats, False, 200)
    model = dict(out.ps_outputs())
    acc = sum(algo.predict(v, model) == y for v, y in data[:50]) / 50
    assert acc >= 0.7, acc
    from fps_b200.models.pa.evaluation import PassiveAggressiveMultiModelEvaluation as EvM
    assert abs(EvM.accuracy(model, data[:50], feats, algo) - 100.0 * acc) < 1e-9
    pred = transformMulticlassWithLongId(list(model.items()))(
        [Right((1000 + i, v)) for i, (v, _) in enumerate(data[:10])], 2, 2, algo, 100, L, feats, False, 200)
    ids = sorted(i for i, _ in pred.worker_outputs())
    assert ids == list(range(1000, 1010))
