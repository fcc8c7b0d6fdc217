"""LEMP top-K: every pruning strategy must return the brute-force top-K (the reference has no test
for its pruning bounds), plus the serving and online learner+generator jobs."""
import random

import numpy as np
import pytest

from fps_b200.api import Left, Right
from fps_b200.models.mf.common import Rating, attachLength
from fps_b200.models.mf.pruning import COORD, INCR, LC, LENGTH, LI, LEMPPruningStrategy
from fps_b200.models.mf.topk import (CollectTopKFromEachWorker, _SortedItems, lemp_topk,
                                     psOnlineLearnerAndGenerator, psTopKGenerator)


def test_strategy_parser():
    assert LEMPPruningStrategy.fromString("length") == LENGTH()
    assert LEMPPruningStrategy.fromString("coord") == COORD()
    assert LEMPPruningStrategy.fromString("incr:3") == INCR(3)
    assert LEMPPruningStrategy.fromString("lc:1.5") == LC(1.5)
    assert LEMPPruningStrategy.fromString("li:5:2.5") == LI(5, 2.5)
    with pytest.raises(ValueError):
        LEMPPruningStrategy.fromString("nope")


@pytest.mark.parametrize("strategy", ["length", "coord", "incr:3", "lc:1.2", "li:3:1.2"])
def test_lemp_equals_brute_force(strategy):
    rng = np.random.default_rng(0)
    k, n = 12, 600
    scale = rng.gamma(2.0, 1.0, size=n)[:, None]
    V = rng.normal(size=(n, k)) * scale
    model = {i: attachLength(V[i]) for i in range(n)}
    items = _SortedItems()
    for i in range(n):
        items.put(i, model[i][0])
    for q in range(20):
        u = attachLength(rng.normal(size=k))
        got = lemp_topk(u, items, model, 10, 50, LEMPPruningStrategy.fromString(strategy)).sorted_desc()
        ref = sorted(((float(V[i] @ u[1]), i) for i in range(n)), reverse=True)[:10]
        assert [i for _, i in got] == [i for _, i in ref]


def test_collect_topk_filters_seen_items():
    c = CollectTopKFromEachWorker(K=2, memory=10, workerParallelism=2)
    r = Rating(1, 7, 1.0, 5)
    assert c.flatMap(Left((r.enrich(0, 100), [(0.9, 7), (0.5, 3)]))) == []
    out = c.flatMap(Left((r.enrich(1, 100), [(0.8, 4), (0.1, 9)])))
    assert out == [(1, 7, 5, [(0.9, 7), (0.8, 4)])]
    r2 = Rating(1, 3, 1.0, 6)
    c.flatMap(Left((r2.enrich(0, 101), [(0.9, 7), (0.5, 3)])))
    out2 = c.flatMap(Left((r2.enrich(1, 101), [(0.8, 4)])))
    assert out2 == [(1, 3, 6, [(0.8, 4), (0.5, 3)])]      # item 7 was seen -> filtered


def test_ps_topk_generator_serving_with_model_load():
    rng = np.random.default_rng(1)
    k, ni, nu = 8, 200, 10
    V = rng.normal(size=(ni, k)); U = rng.normal(size=(nu, k))
    model = [Left((i, attachLength(V[i]))) for i in range(ni)] + [Right((u, attachLength(U[u]))) for u in range(nu)]
    ratings = [Rating(u, 0, 1.0, u) for u in range(nu)] + [Rating(999, 0, 1.0, 99)]
    out = psTopKGenerator(ratings, model, numFactors=k, K=5, workerK=5, bucketSize=20, pruningAlgorithm=LENGTH(),
                          pullLimit=50, workerParallelism=3, psParallelism=2, iterationWaitTime=250)
    by_ts = {ts: topk for _, ts, topk in out}
    assert by_ts[99] == []                             # unknown user -> empty list
    for u in range(nu):
        ref = list(np.argsort(-(V @ U[u]))[:5])
        assert [i for _, i in by_ts[u]] == ref


def test_online_learner_and_generator_emits_one_topk_per_rating():
    r = random.Random(5)
    ratings = [Rating(r.randrange(15), r.randrange(30), 1.0, t) for t in range(150)]
    out = psOnlineLearnerAndGenerator(ratings, numFactors=6, rangeMin=-0.1, rangeMax=0.1, learningRate=0.1,
                                      negativeSampleRate=2, K=5, workerK=5, bucketSize=8, pullLimit=20,
                                      workerParallelism=3, psParallelism=2, iterationWaitTime=300, seed=2)
    assert len(out) == len(ratings)
    assert all(len(topk) <= 5 for *_x, topk in out)
    assert any(len(topk) > 0 for *_x, topk in out[50:])


def test_device_api_seen_filter_matches_collect_topk_semantics():
    """`_seen_filter` (used by the device backends) == the sequential part of CollectTopKFromEachWorker:
    drop items the user rated within the last `memory` events, then remember the rated item."""
    from fps_b200.api import Left
    from fps_b200.models.mf.common import Rating
    from fps_b200.models.mf.device_api import _seen_filter
    from fps_b200.models.mf.topk import CollectTopKFromEachWorker

    cand = [(float(10 - i), i) for i in range(10)]                       # items 0..9, best first
    events = [Rating(1, 0, 1.0, 0), Rating(1, 1, 1.0, 1), Rating(2, 0, 1.0, 2), Rating(1, 2, 1.0, 3),
              Rating(1, 3, 1.0, 4)]
    for memory in (0, 1, 2, -1):
        rows = [(e.user, e.item, e.timestamp, list(cand)) for e in events]
        got = _seen_filter(rows, 4, memory)
        ref = CollectTopKFromEachWorker(4, memory, 1)
        want = []
        for rid, e in enumerate(events):
            want += ref.flatMap(Left((e.enrich(0, rid), list(cand))))
        assert got == want, memory
    assert [i for _, i in _seen_filter([(1, 0, 0, list(cand)), (1, 5, 1, list(cand))], 3, 5)[1][3]] == [1, 2, 3]
