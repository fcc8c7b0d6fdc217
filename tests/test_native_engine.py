"""Native host engine (C++ threads + SPSC rings) for asynchronous MF: exact in sequential mode, passes the
reference's quality gate (T/matrix/factorization/PSOfflineMatrixFactorizationTest.scala:55-103) in parallel."""
import random

import numpy as np
import pytest
import torch

from fps_b200.models.mf.common import Rating
from fps_b200.models.mf.offline import psOfflineMF
from fps_b200.models.mf.online import psOnlineMF
from fps_b200.ops import host

M64 = (1 << 64) - 1


def _splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
    return x ^ (x >> 31)


def _init(seed, id_, j, lo, hi):
    h = _splitmix64(seed ^ _splitmix64((id_ * 0x100000001B3 + j) & M64))
    return np.float32(lo) + np.float32(hi - lo) * np.float32((h >> 40) * (1.0 / 16777216.0))


def test_sequential_mode_equals_a_plain_sgd_loop():
    """1 worker, 1 server, pullLimit 1 => no concurrency => bit-for-bit the textbook loop."""
    rnd = random.Random(0)
    nu, ni, k, n = 12, 9, 6, 200
    u = [rnd.randrange(nu) for _ in range(n)]
    i = [rnd.randrange(ni) for _ in range(n)]
    r = [rnd.random() for _ in range(n)]
    for plain in (False, True):
        ut, it, utc, itc, sse = host.mf_train(torch.tensor(u), torch.tensor(i), torch.tensor(r), nu, ni, k, -0.1, 0.1,
                                             0.05, workers=1, servers=1, pull_limit=1, epochs=2, seed=3,
                                             plain_residual=plain)
        U = np.array([[_init(3 * 2 + 2, a, j, -0.1, 0.1) for j in range(k)] for a in range(nu)], dtype=np.float32)
        V = np.array([[_init(3 * 2 + 1, a, j, -0.1, 0.1) for j in range(k)] for a in range(ni)], dtype=np.float32)
        for _ep in range(2):
            for a, b, c in zip(u, i, r):
                v = V[b].copy()
                dot = np.float32(0)
                for j in range(k):
                    dot = np.float32(dot + U[a, j] * v[j])
                resid = np.float32(np.float32(c) - dot)
                e = resid if plain else np.float32(1.0) / (np.float32(1.0) + np.exp(-resid, dtype=np.float32))
                g = np.float32(np.float32(0.05) * e)
                delta = (g * U[a]).astype(np.float32)
                U[a] = (U[a] + g * v).astype(np.float32)
                V[b] = (V[b] + delta).astype(np.float32)
        np.testing.assert_allclose(ut, U, rtol=2e-6, atol=1e-7)
        np.testing.assert_allclose(it, V, rtol=2e-6, atol=1e-7)
        assert utc.sum() == len(set(u)) and itc.sum() == len(set(i))


def test_parallel_engine_passes_the_reference_quality_gate():
    rnd = random.Random(47)
    seen, ratings = set(), []
    while len(ratings) < 100:
        a, b = rnd.randrange(20), rnd.randrange(15)
        if (a, b) not in seen:
            seen.add((a, b))
            ratings.append(Rating(a, b, rnd.random(), 0))
    out = psOfflineMF(ratings, numFactors=15, rangeMin=0.0, rangeMax=0.3, learningRate=0.05, iterations=25,
                      pullLimit=10, workerParallelism=4, psParallelism=4, seed=1, backend="native",
                      plain_residual=True)             # same configuration as the host-tier gate (test_mf_host.py)
    U = dict(out.worker_outputs()); V = dict(out.ps_outputs())
    assert set(U) == {r.user for r in ratings} and set(V) == {r.item for r in ratings}
    rmse = np.sqrt(np.mean([(float(np.dot(U[r.user], V[r.item])) - r.rating) ** 2 for r in ratings]))
    assert rmse <= 0.5, rmse


def test_online_backend_native_learns_and_validates_arguments():
    g = torch.Generator().manual_seed(1)
    nu, ni, n = 300, 100, 60000
    P, Q = torch.rand(nu, 3, generator=g), torch.rand(ni, 3, generator=g)
    u = torch.randint(0, nu, (n,), generator=g); i = torch.randint(0, ni, (n,), generator=g)
    r = (P[u] * Q[i]).sum(1) / 3
    recs = [Rating(int(a), int(b), float(c), 0) for a, b, c in zip(u, i, r)]
    out = psOnlineMF(recs, numFactors=8, rangeMin=-0.1, rangeMax=0.1, learningRate=0.05, pullLimit=32,
                     workerParallelism=3, psParallelism=2, seed=2, backend="native", plain_residual=True)
    U = dict(out.worker_outputs()); V = dict(out.ps_outputs())
    pred = np.array([np.dot(U[int(a)], V[int(b)]) for a, b in zip(u[-5000:], i[-5000:])])
    assert np.sqrt(np.mean((pred - r[-5000:].numpy()) ** 2)) < 0.12
    with pytest.raises(ValueError):
        host.mf_train(torch.tensor([5]), torch.tensor([0]), torch.tensor([1.0]), 3, 3)      # user id out of range
    with pytest.raises(ValueError):
        host.mf_train(torch.tensor([0]), torch.tensor([0]), torch.tensor([1.0]), 3, 3, num_factors=500)


def test_native_engine_is_race_free_under_thread_sanitizer(tmp_path):
    """Race detection (absent in the reference): randomised worker / server / pull-limit
    configurations of the native engine under ThreadSanitizer."""
    import os
    import shutil
    import subprocess

    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    probe = subprocess.run(["g++", "-fsanitize=thread", "-x", "c++", "-", "-o", str(tmp_path / "probe")],
                           input="int main(){return 0;}", text=True, capture_output=True)
    if probe.returncode != 0:
        pytest.skip("g++ has no ThreadSanitizer runtime")
    r = subprocess.run(["bash", os.path.join(repo, "scripts", "tsan_host.sh")], cwd=repo, capture_output=True,
                       text=True, timeout=900, env={**os.environ, "TMPDIR": str(tmp_path)})
    assert r.returncode == 0 and "tsan: clean" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def _sparse(rnd, feats, nnz):
    from fps_b200.models.pa.sparse import SparseVector

    idx = rnd.sample(range(feats), nnz)
    return SparseVector(idx, [rnd.gauss(0, 1) for _ in idx], feats)


@pytest.mark.parametrize("algo_name,C", [("PA", 0.0), ("PAI", 0.05), ("PAII", 0.5)])
@pytest.mark.parametrize("range_part", [False, True])
def test_native_pa_sequential_equals_host_algorithm(algo_name, C, range_part):
    """1 worker => examples are processed one after the other => must equal the host algorithm step by step
    (any number of server threads: per-worker FIFO rings order pushes before later pulls)."""
    from fps_b200.api import Left
    from fps_b200.models.pa.algorithms import PassiveAggressiveBinaryAlgorithm as B
    from fps_b200.models.pa.ps import transformBinary

    rnd = random.Random(1)
    feats = 3000
    algo = {"PA": B.buildPA(), "PAI": B.buildPAI(C), "PAII": B.buildPAII(C)}[algo_name]
    data, w = [], {}
    for _ in range(80):
        v, y = _sparse(rnd, feats, 40), rnd.random() < 0.5
        data.append(Left((v, y)))
        for i, d in algo.delta(v, {i: w.get(i, 0.0) for i in v.indices.tolist()}, y):
            w[i] = w.get(i, 0.0) + d
    out = transformBinary()(data, 1, 3, algo, 7, feats, range_part, 100, backend="native")
    got = dict(out.ps_outputs())
    assert {i for i, x in got.items() if x != 0} == {i for i, x in w.items() if x != 0}
    for i, x in w.items():
        assert abs(got[i] - x) < 1e-4 * max(1.0, abs(x))


def test_native_pa_accuracy_gate_and_predict_with_model_load():
    """PassiveAggressiveParameterServerTest.scala:44-100 on the native engine: accuracy >= 80 %."""
    from fps_b200.api import Left, Right
    from fps_b200.models.pa.algorithms import PassiveAggressiveBinaryAlgorithm as B
    from fps_b200.models.pa.ps import transformBinary

    rnd = random.Random(7)
    feats = 50_000
    w_true, data = {}, []
    for _ in range(80):
        v = _sparse(rnd, feats, 500)
        s = sum(x * w_true.setdefault(i, rnd.gauss(0, 1)) for i, x in v.activeIterator())
        data.append((v, s > 0))
    out = transformBinary()([Left(d) for d in data] * 3, 3, 3, B.buildPA(), 100, feats, True, 100, backend="native")
    model = dict(out.ps_outputs())
    assert sum((v.dot(model) > 0) == y for v, y in data[:20]) / 20 >= 0.8
    pred = transformBinary(list(model.items()))([Right((i, v)) for i, (v, _) in enumerate(data[:20])], 3, 3,
                                                B.buildPA(), 100, feats, True, 100, backend="native")
    got = pred.worker_outputs()
    assert len(got) == 20 and sum(bool(p) == y for (_v, p), (_v2, y) in zip(got, data[:20])) / 20 >= 0.8


@pytest.mark.parametrize("which", ["OVA_PA", "OVA_PAI", "OVA_PAII", "PB", "ML"])
def test_native_multiclass_pa_sequential_equals_host_algorithm(which):
    from fps_b200.api import Left, Right
    from fps_b200.models.pa.algorithms import PassiveAggressiveCostBased as CB
    from fps_b200.models.pa.algorithms import PassiveAggressiveOneVersusAll as OVA
    from fps_b200.models.pa.ps import transformMulticlass

    rnd = random.Random(2)
    feats, L = 500, 5
    cost = lambda a, b: 0.0 if a == b else 1.0 + 0.25 * abs(a - b)      # noqa: E731
    algo = {"OVA_PA": OVA.buildPA(L), "OVA_PAI": OVA.buildPAI(L, 0.1), "OVA_PAII": OVA.buildPAII(L, 0.5),
            "PB": CB.buildPB(cost, L), "ML": CB.buildML(cost, L)}[which]
    data, w, host_pred = [], {}, []
    for step in range(60):
        v, y = _sparse(rnd, feats, 25), rnd.randrange(L)
        model = {i: w.get(i, np.zeros(L)) for i in v.indices.tolist()}
        if step % 4 == 3:                                   # every 4th record is a query
            data.append(Right((step, v)))
            host_pred.append(algo.predict(v, model))
            continue
        data.append(Left((v, y)))
        for i, d in algo.delta(v, model, y):
            w[i] = w.get(i, np.zeros(L)) + d
    out = transformMulticlass()(data, 1, 3, algo, 7, L, feats, False, 100, backend="native")
    assert [p for (_v, p) in out.worker_outputs()] == host_pred
    got = dict(out.ps_outputs())
    for i, x in w.items():
        np.testing.assert_allclose(got[i], x, rtol=1e-4, atol=1e-5)


def test_native_mf_negative_sampling_matches_a_python_replica():
    """Sequential mode again, now with negativeSampleRate / userMemory: the sampler (items seen so far by the
    worker minus the user's recent items, splitmix64 stream) is replicated in Python, so tables must agree."""
    rnd = random.Random(3)
    nu, ni, k, n, neg, mem, seed = 15, 12, 5, 150, 2, 3, 9
    u = [rnd.randrange(nu) for _ in range(n)]
    i = [rnd.randrange(ni) for _ in range(n)]
    r = [1.0] * n
    ut, it, utc, itc, sse = host.mf_train(torch.tensor(u), torch.tensor(i), torch.tensor(r), nu, ni, k, -0.1, 0.1,
                                         0.05, workers=1, servers=1, pull_limit=1, epochs=1, seed=seed,
                                         plain_residual=True, negative_sample_rate=neg, user_memory=mem)
    U = np.array([[_init(seed * 2 + 2, a, j, -0.1, 0.1) for j in range(k)] for a in range(nu)], dtype=np.float32)
    V = np.array([[_init(seed * 2 + 1, a, j, -0.1, 0.1) for j in range(k)] for a in range(ni)], dtype=np.float32)
    rng = _splitmix64(seed ^ (0xA5A5A5A5 + 0))
    item_ids, known, seen = [], set(), {}
    n_updates = 0

    def update(a, b, c):
        nonlocal n_updates
        v = V[b].copy()
        dot = np.float32(0)
        for j in range(k):
            dot = np.float32(dot + U[a, j] * v[j])
        g = np.float32(np.float32(0.05) * np.float32(np.float32(c) - dot))
        delta = (g * U[a]).astype(np.float32)
        U[a] = (U[a] + g * v).astype(np.float32)
        V[b] = (V[b] + delta).astype(np.float32)
        n_updates += 1

    for a, b, c in zip(u, i, r):
        q = seen.setdefault(a, [])
        if len(q) >= mem and q:
            q.pop(0)
        q.append(b)
        for _ in range(max(0, min(len(item_ids) - len(q), neg))):
            while True:
                rng = _splitmix64(rng)
                cand = item_ids[rng % len(item_ids)]
                if cand not in q:
                    break
            update(a, cand, 0.0)
        if b not in known:
            known.add(b); item_ids.append(b)
        update(a, b, c)
    assert n_updates > n * 2                     # negatives were actually drawn
    np.testing.assert_allclose(ut, U, rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(it, V, rtol=2e-6, atol=1e-7)
