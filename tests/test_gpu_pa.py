"""Fused passive-aggressive kernel vs the host-tier algorithms (the semantic oracle)."""
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _vec(r, feats, nnz):
    from fps_b200.models.pa.sparse import SparseVector

    idx = r.sample(range(feats), nnz)
    return SparseVector(idx, [r.gauss(0, 1) for _ in idx], feats)


@pytest.mark.parametrize("algo_name,C", [("PA", 0.0), ("PAI", 0.05), ("PAII", 0.5)])
@pytest.mark.parametrize("range_part", [False, True])
def test_binary_sequential_matches_host_algorithm(algo_name, C, range_part):
    """batch_size=1 => no intra-batch races => must equal the host algorithm step by step."""
    from fps_b200.models.pa.algorithms import PassiveAggressiveBinaryAlgorithm as B
    from fps_b200.models.pa.device import DevicePassiveAggressive

    torch.cuda.set_device(0)
    r = random.Random(1)
    feats = 3000
    host = {"PA": B.buildPA(), "PAI": B.buildPAI(C), "PAII": B.buildPAII(C)}[algo_name]
    pa = DevicePassiveAggressive(feats, 1, True, algo_name, C, None, range_part)
    w = {}
    for _ in range(60):
        v, y = _vec(r, feats, 40), r.random() < 0.5
        pred_host = host.predict(v, w)
        for i, d in host.delta(v, {i: w.get(i, 0.0) for i in v.indices.tolist()}, y):
            w[i] = w.get(i, 0.0) + d
        pred_dev = pa.step([v], [1 if y else -1])[0]
        assert bool(pred_dev) == pred_host
    dev = dict(pa.model())
    assert set(dev) == {i for i, x in w.items() if x != 0}
    for i, x in dev.items():
        assert abs(x - w[i]) < 1e-4 * max(1, abs(w[i]))
    pa.close()


@pytest.mark.parametrize("L,block_kernel", [(5, False), (3, False), (37, False), (5, True)])
@pytest.mark.parametrize("which", ["OVA_PA", "OVA_PAI", "PB", "ML"])
def test_multiclass_sequential_matches_host_algorithm(which, L, block_kernel):
    """warp-per-example kernel with 2 / 1 / 16 lanes per row, and the block-per-example kernel."""
    from fps_b200.ops import native

    native.lib().fps_set_pa_variant(1 if block_kernel else 0)
    try:
        _multiclass_case(which, L)
    finally:
        native.lib().fps_set_pa_variant(0)


def _multiclass_case(which, L):
    from fps_b200.models.pa.algorithms import PassiveAggressiveCostBased as CB
    from fps_b200.models.pa.algorithms import PassiveAggressiveOneVersusAll as OVA
    from fps_b200.models.pa.device import DevicePassiveAggressive, algo_to_device

    torch.cuda.set_device(0)
    r = random.Random(2)
    feats = 500
    cost = lambda a, b: 0.0 if a == b else 1.0 + 0.25 * abs(a - b)
    host = {"OVA_PA": OVA.buildPA(L), "OVA_PAI": OVA.buildPAI(L, 0.1), "PB": CB.buildPB(cost, L),
            "ML": CB.buildML(cost, L)}[which]
    name, C, cm = algo_to_device(host)
    pa = DevicePassiveAggressive(feats, L, False, name, C, cm)
    w = {}
    for _ in range(50):
        v, y = _vec(r, feats, 25), r.randrange(L)
        model = {i: w.get(i, np.zeros(L)) for i in v.indices.tolist()}
        pred_host = host.predict(v, model)
        for i, d in host.delta(v, model, y):
            w[i] = w.get(i, np.zeros(L)) + d
        assert pa.step([v], [y])[0] == pred_host
    dev = dict(pa.model())
    for i, x in dev.items():
        np.testing.assert_allclose(x, w[i], rtol=1e-4, atol=1e-5)
    pa.close()


def test_binary_accuracy_gate_batched_with_predict_and_model_load():
    """Quality gate of PassiveAggressiveParameterServerTest.scala:44-100 through the public API."""
    from fps_b200.api import Left, Right
    from fps_b200.models.pa.algorithms import PassiveAggressiveBinaryAlgorithm as B
    from fps_b200.models.pa.ps import transformBinary

    torch.cuda.set_device(0)
    r = random.Random(7)
    feats = 500_000
    w_true = {}
    data = []
    for _ in range(80):
        v = _vec(r, feats, 2000)
        s = sum(x * w_true.setdefault(i, r.gauss(0, 1)) for i, x in v.activeIterator())
        data.append((v, s > 0))
    out = transformBinary()([Left(d) for d in data] * 3, 3, 3, B.buildPA(), 10000, feats, True, 100,
                            backend="device", batch_size=8)
    model = dict(out.ps_outputs())
    acc = sum((v.dot(model) > 0) == y for v, y in data[:20]) / 20
    assert acc >= 0.8, acc
    pred = transformBinary(list(model.items()))([Right((i, v)) for i, (v, _) in enumerate(data[:20])], 3, 3,
                                                B.buildPA(), 10000, feats, True, 100, backend="device")
    got = {v: p for v, p in pred.worker_outputs()}
    assert sum(bool(got[v]) == y for v, y in data[:20]) / 20 >= 0.8
