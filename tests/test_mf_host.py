"""Host-tier matrix factorisation: the reference's model-quality gate
(T/matrix/factorization/PSOfflineMatrixFactorizationTest.scala:55-103: RMSE <= 0.5)."""
import random

import numpy as np

from fps_b200.models.mf.common import (PseudoRandomFactorInitializerDescriptor, Rating, SGDUpdater,
                                       TopKQueue, attachLength, vectorSum, FactorIsNotANumberException)
from fps_b200.models.mf.offline import psOfflineMF
from fps_b200.models.mf.online import psOnlineMF


def _ratings(seed=47, n=100, nu=20, ni=15):
    r = random.Random(seed)
    seen, out = set(), []
    for _ in range(n):
        u, i = r.randrange(nu), r.randrange(ni)
        if (u, i) not in seen:
            seen.add((u, i))
            out.append(Rating(u, i, r.random()))
    return out


def _rmse(out, ratings):
    users, items = {}, {}
    for u, v in out.worker_outputs():
        users[u] = v
    for i, v in out.ps_outputs():
        items[i] = v
    se = [(r.rating - float(np.dot(users[r.user], items[r.item]))) ** 2 for r in ratings]
    return (sum(se) / len(se)) ** 0.5


def test_offline_mf_rmse_gate_4x4():
    ratings = _ratings()
    out = psOfflineMF(ratings, numFactors=15, rangeMin=0.0, rangeMax=0.3, learningRate=0.05,
                      iterations=25, pullLimit=10, workerParallelism=4, psParallelism=4,
                      iterationWaitTime=300, seed=1, plain_residual=True)
    assert _rmse(out, ratings) <= 0.5


def test_online_mf_runs_and_negative_samples():
    ratings = _ratings(n=200) * 1
    out = psOnlineMF(ratings, numFactors=8, learningRate=0.05, negativeSampleRate=2, userMemory=4,
                     pullLimit=5, workerParallelism=3, psParallelism=2, iterationWaitTime=150, seed=3)
    n_pos = len(ratings)
    assert len(out.worker_outputs()) >= n_pos           # positives + sampled negatives
    assert len(out.ps_outputs()) == len(out.worker_outputs())   # SimplePSLogic emits every push
    users = {u for u, _ in out.worker_outputs()}
    assert users == {r.user for r in ratings}


def test_sgd_updater_parity_and_helpers():
    u, v = np.array([1.0, 2.0]), np.array([0.5, -1.0])
    du, dv = SGDUpdater(0.1).delta(1.0, u, v)
    e = 1 / (1 + np.exp(-(1.0 - (0.5 - 2.0))))
    np.testing.assert_allclose(du, 0.1 * e * v); np.testing.assert_allclose(dv, 0.1 * e * u)
    a = PseudoRandomFactorInitializerDescriptor(4).open().nextFactor(7)
    b = PseudoRandomFactorInitializerDescriptor(4).open().nextFactor(7)
    np.testing.assert_array_equal(a, b)
    assert attachLength(np.array([3.0, 4.0]))[0] == 5.0
    q = TopKQueue(2)
    for s, i in [(0.1, 1), (0.9, 2), (0.5, 3)]:
        q.push(s, i)
    assert q.sorted_desc() == [(0.9, 2), (0.5, 3)]
    try:
        vectorSum(np.array([np.nan]), np.array([1.0]))
        assert False
    except FactorIsNotANumberException:
        pass
