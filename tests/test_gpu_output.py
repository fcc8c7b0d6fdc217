"""E5 worker output stream on the device tier + its device-side count / timer flush (K11):
PSOnlineMatrixFactorizationWorker.scala:52 (``ps.output((user, userVector))`` after every update),
CountLogic.scala:5-29, TimerLogic.scala:6-51."""
import time

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_per_update_outputs_equal_the_sequential_oracle():
    from fps_b200.models.mf.common import Rating
    from fps_b200.models.mf.online import psOnlineMF
    from tests.philox_ref import init_rows_ref

    torch.cuda.set_device(0)
    rng = np.random.RandomState(1)
    nu, ni, k, lr, seed = 17, 11, 10, 0.05, 3
    ratings = [Rating(int(rng.randint(nu)), int(rng.randint(ni)), float(rng.rand()), t) for t in range(150)]
    out = psOnlineMF(ratings, numFactors=k, rangeMin=-0.3, rangeMax=0.3, learningRate=lr, seed=seed,
                     backend="device", batch_size=1, updateOutput=1, numUsers=nu, numItems=ni)
    got = [x.value for x in out.collect() if x.is_left]
    # single-stream oracle: same Philox init by id, reference update rule e = sigmoid(r - u.v)
    U = init_rows_ref(np.arange(nu), k, seed * 2 + 2, -0.3, 0.3)[:, :k].astype(np.float64)
    V = init_rows_ref(np.arange(ni), k, seed * 2 + 1, -0.3, 0.3)[:, :k].astype(np.float64)
    want = []
    for r in ratings:
        u, v = U[r.user].copy(), V[r.item].copy()
        e = 1.0 / (1.0 + np.exp(-(r.rating - u @ v)))
        U[r.user] = u + lr * e * v
        V[r.item] = v + lr * e * u
        want.append((r.user, U[r.user].copy()))
    assert [g[0] for g in got] == [w[0] for w in want]            # one record per update, in stream order
    for (gu, gv), (wu, wv) in zip(got, want):
        np.testing.assert_allclose(gv, wv, rtol=2e-5, atol=2e-6)
    items = dict(x.value for x in out.collect() if x.is_right)
    for i in range(ni):
        np.testing.assert_allclose(items[i], V[i], rtol=2e-5, atol=2e-6)
    c = out.output_ring.counters()
    assert c["published"] == len(ratings) and c["dropped"] == 0
    out.model.close()


def _model_with_ring(**ring_kw):
    from fps_b200.models.mf.device import DeviceOnlineMF
    from fps_b200.runtime.output_ring import OutputRing

    dev = torch.device("cuda", 0)
    ring = OutputRing(16, dev, **ring_kw)
    m = DeviceOnlineMF(1000, 500, 16, learning_rate=0.01, seed=1, output_ring=ring)
    return m, ring, dev


def _batch(dev, n, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randint(0, 1000, (n,), generator=g, dtype=torch.int32).to(dev),
            torch.randint(0, 500, (n,), generator=g, dtype=torch.int32).to(dev),
            torch.rand(n, generator=g).to(dev))


def test_output_count_and_timer_flush_are_decided_on_the_device():
    m, ring, dev = _model_with_ring(flush_count=50, ring_capacity=4096, staging_capacity=4096)
    for s in range(2):
        m.step(*_batch(dev, 20, s))
    torch.cuda.synchronize()
    assert ring.poll()[0].size == 0 and ring.counters()["staged"] == 40      # count not reached: nothing published
    m.step(*_batch(dev, 20, 2))
    torch.cuda.synchronize()
    ids, vecs = ring.poll()
    assert ids.size == 60 and vecs.shape == (60, 16) and ring.counters()["staged"] == 0
    m.close()
    m, ring, dev = _model_with_ring(flush_count=0, flush_interval_ms=20, ring_capacity=4096, staging_capacity=4096)
    m.step(*_batch(dev, 20, 0)); torch.cuda.synchronize()
    assert ring.poll()[0].size == 0                                          # deadline not passed
    time.sleep(0.03)
    m.step(*_batch(dev, 20, 1)); torch.cuda.synchronize()
    assert ring.poll()[0].size == 40                                         # globaltimer deadline passed
    m.close()
    m, ring, dev = _model_with_ring(flush_count=30, flush_interval_ms=20, require="all", ring_capacity=4096,
                                    staging_capacity=4096)
    m.step(*_batch(dev, 40, 0)); torch.cuda.synchronize()
    assert ring.poll()[0].size == 0                                          # count reached, timer not (AND)
    time.sleep(0.03)
    m.step(*_batch(dev, 1, 1)); torch.cuda.synchronize()
    assert ring.poll()[0].size == 41
    m.close()


def test_output_sampling_and_full_ring_accounting():
    m, ring, dev = _model_with_ring(every=4, flush_count=1, ring_capacity=4096, staging_capacity=4096)
    u, i, r = _batch(dev, 1001, 5)
    m.step(u, i, r); torch.cuda.synchronize()
    ids, vecs = ring.poll()
    assert ids.size == 251                                                    # one update in four
    assert set(ids.tolist()) <= set(u.cpu().tolist())
    # emitted vector = the updated row as this update saw it: close to the final row for rarely-hit users
    m.close()
    m, ring, dev = _model_with_ring(flush_count=1, ring_capacity=64, staging_capacity=4096)
    m.step(*_batch(dev, 200, 6)); torch.cuda.synchronize()
    assert ring.poll()[0].size == 64 and ring.counters()["dropped"] == 136   # nobody read: newest dropped, counted
    m.step(*_batch(dev, 10, 7)); torch.cuda.synchronize()
    assert ring.poll()[0].size == 10                                          # the reader freed the slots
    m.close()
