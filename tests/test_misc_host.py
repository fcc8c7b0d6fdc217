"""Offline PA pipeline, experiment CLIs, model IO, metrics, rate-replay source."""
import random

import pytest

from fps_b200.models.pa.offline import PassiveAggressiveFilter, paBinaryClassificationOffline
from fps_b200.models.pa.sparse import LegacySparseVector
from fps_b200.models.sketch import experiments
from fps_b200.models.sketch.hashing import java_string_hash
from fps_b200.utils import InputSource, metrics, model_io


def test_legacy_pa_filter_variants():
    x = LegacySparseVector({0: 1.0, 2: 2.0})
    w = {0: 0.0, 2: 0.0}
    assert PassiveAggressiveFilter.buildPAF().delta(x, w, 1) == {0: 0.2, 2: 0.4}
    assert PassiveAggressiveFilter.buildPAFI(0.1).delta(x, w, -1) == {0: -0.1, 2: -0.2}
    assert PassiveAggressiveFilter.buildPAF().predict(x, {0: 1.0, 2: -0.1}) == 1


def test_pa_offline_train_then_predict():
    r = random.Random(4)
    feats = 60
    w_true = [r.gauss(0, 1) for _ in range(feats)]

    def vec():
        idx = r.sample(range(feats), 12)
        return LegacySparseVector({i: r.gauss(0, 1) for i in idx})

    label = lambda v: 1 if sum(w_true[k] * x for k, x in v.getValues().items()) > 0 else -1
    train = [(v, label(v)) for v in (vec() for _ in range(120))]
    test = [v for v, _ in train[:30]]
    out = paBinaryClassificationOffline(train, test, 3, 2, iterations=5, pafType=0, pullLimit=50,
                                        iterationWaitTime=400, seed=1)
    preds = out.worker_outputs()
    assert len(preds) == 30
    acc = sum(p == label(v) for v, p in preds) / 30
    assert acc >= 0.8, acc
    assert len(out.ps_outputs()) > 0


def test_experiment_mains_roundtrip(tmp_path):
    words = ["cat", "dog", "fish", "bird"]
    (tmp_path / "words.txt").write_text("\n".join(words))
    (tmp_path / "q.txt").write_text("cat\n")
    lines = []
    for i in range(60):
        text = "cat dog" if i % 3 else ("bird fish" if i % 2 else "cat fish")
        lines.append(f"{1000 + i}|{3600 * (i % 2)}|a|b|c|{text}")
    (tmp_path / "tweets.txt").write_text("\n".join(lines))
    t, w, q = str(tmp_path / "tweets.txt"), str(tmp_path / "words.txt"), str(tmp_path / "q.txt")
    m, p = str(tmp_path / "bloom.model"), str(tmp_path / "bloom.pred")
    assert experiments.main(["BloomFilterExp", t, w, "|", m, "2", "2", "150", "3", "256"]) == 0
    keys = {int(l.split(":")[0]) for l in open(m)}
    assert keys == {java_string_hash(x) for x in words}
    assert experiments.main(["BloomFilterPredictExp", m, w, q, p, "2", "2", "150", "10", "3", "256", "2"]) == 0
    line = open(p).read().strip()
    assert line.startswith("cat - ") and "dog" in line
    m2, p2 = str(tmp_path / "tow.model"), str(tmp_path / "tow.pred")
    assert experiments.main(["TugOfWarExp", t, w, "|", m2, "2", "2", "150", "64"]) == 0
    assert experiments.main(["TugOfWarPredictExp", m2, w, q, p2, "2", "2", "150", "10", "64", "4", "2"]) == 0
    assert open(p2).read().startswith("cat - ")
    m3, p3 = str(tmp_path / "mh.model"), str(tmp_path / "mh.pred")
    assert experiments.main(["MinHashExp", t, w, "|", m3, "2", "2", "150", "32"]) == 0
    assert experiments.main(["MinHashPredictExp", m3, w, q, t, "|", p3, "2", "2", "150", "10", "32", "3"]) == 0
    assert open(p3).read().startswith("cat - ")
    m4 = str(tmp_path / "tab.model")
    assert experiments.main(["TimeAwareBloomFilterExp", t, w, "|", m4, "2", "2", "150", "3", "256", "0", "1"]) == 0
    assert all(";" in l.split(":")[0] for l in open(m4))
    assert experiments.main(["nope"]) == 2


def test_model_io_and_metrics(tmp_path):
    path = str(tmp_path / "m.txt")
    model_io.write_text(path, [(3, [0.5, 1.5]), (7, [2.0, 4.0])])
    assert list(model_io.read_text(path)) == [(3, [0.5, 1.5]), (7, [2.0, 4.0])]
    model_io.write_npz(str(tmp_path / "m.npz"), [1, 2], [[1.0], [2.0]])
    ids, vals = model_io.read_npz(str(tmp_path / "m.npz"))
    assert ids.tolist() == [1, 2] and vals.shape == (2, 1)
    assert abs(metrics.rmse([(1, 0), (0, 1)]) - 1.0) < 1e-12 and metrics.accuracy([(1, 1), (0, 1)]) == 0.5
    sink = metrics.NDCGSink(3, period_of=lambda t: t // 10)
    sink.invoke((1, 5, 3, [(0.9, 5), (0.1, 2)])); sink.invoke((1, 9, 4, [(0.9, 5), (0.5, 9)])); sink.invoke((2, 1, 15, []))
    res = sink.result()
    assert res[0][0] == 0 and abs(res[0][1] - (1 + 1 / 1.584962500721156) / 2) < 1e-9 and res[0][2] == 1.0
    assert res[1] == (1, 0.0, 0.0, 1)


def test_input_source_rate_replay():
    clock = {"t": 100.0}
    slept = []

    def sleep(s):
        slept.append(s); clock["t"] += s

    events = [(0, "a"), (1000, "b"), (3000, "c")]
    src = InputSource(None, servingSpeed=10, events=events, event_time=lambda e: e[0], clock=lambda: clock["t"], sleep=sleep)
    assert [e[1] for e in src] == ["a", "b", "c"]
    assert abs(sum(slept) - 0.3) < 1e-6            # 3000 ms of data at 10x = 0.3 s
    src2 = InputSource(None, 1, events=events, event_time=lambda e: e[0])
    src2.cancel()
    assert list(src2) == []


def test_counters_prometheus_text():
    from fps_b200.utils.metrics import Counters

    c = Counters()
    c.inc("ps_pull_rows", 10); c.inc("ps_pull_rows", 5); c.set("ring occupancy/max", 3)
    text = c.prometheus_text()
    assert "fps_b200_ps_pull_rows 15\n" in text and "fps_b200_ring_occupancy_max 3\n" in text


# ---- round-2 host-side fixes -----------------------------------------------------------------------------------
def test_batch_triggers_are_cleared_when_the_buffer_is_swapped_out():
    """After a flush the next batch is measured from zero (a stale count made batches drift)."""
    from fps_b200.protocol.combination import CombinationLogic, CountLogic, TimerLogic, any_of

    got = []
    lg = CombinationLogic(any_of, [CountLogic(3), TimerLogic(3600.0)])
    for i in range(2):
        lg.logic(lambda data, i=i: data.append(i), got.append)
    assert lg.flush(got.append) is True and got == [[0, 1]]          # idle flush with 2 of 3 buffered
    assert lg.flush(got.append) is False                               # nothing left: reports "emitted nothing"
    for i in range(3):
        lg.logic(lambda data, i=i: data.append(10 + i), got.append)
    assert got == [[0, 1], [10, 11, 12]]                               # a full batch of 3, not 1
    assert lg.combinables[1].containsData is False
    lg.close()


def test_seeded_mf_workers_get_their_own_random_streams():
    from fps_b200.models.mf.online import PSOnlineMatrixFactorizationWorker
    from fps_b200.runtime.local_engine import assign_subtask, clone_logic
    from fps_b200.limiter import addPullLimiter

    proto = addPullLimiter(PSOnlineMatrixFactorizationWorker(4, -1.0, 1.0, 0.1, 8, 0, seed=7), 10)
    copies = [clone_logic(proto) for _ in range(3)]
    for i, c in enumerate(copies):
        assign_subtask(c, i, 3)
        c.open()
    firsts = [tuple(c.inner._factor_init().nextFactor(0)) for c in copies]
    assert len(set(firsts)) == 3                                       # identical seeds used to give identical vectors
    again = clone_logic(proto); assign_subtask(again, 1, 3); again.open()
    assert tuple(again.inner._factor_init().nextFactor(0)) == firsts[1]   # still deterministic per subtask


def test_native_pa_rejects_malformed_csr():
    import numpy as np
    from fps_b200.ops import host

    with pytest.raises(ValueError):
        host.pa_binary([0, 2], [0, 1], [1.0, 1.0], [1, -1], 4)        # row_ptr needs n + 1 entries
    with pytest.raises(ValueError):
        host.pa_binary([0, 1, 3], [0, 1], [1.0, 1.0], [1, -1], 4)     # row_ptr[-1] != nnz


def test_synthetic_low_rank_ratings_are_a_pure_function_of_the_ids():
    import torch
    from fps_b200.utils.synthetic import lowrank_ratings

    u = torch.randint(0, 10_000_000, (50_000,)); i = torch.randint(0, 1_000_000, (50_000,))
    r = lowrank_ratings(u, i)
    assert torch.equal(r, lowrank_ratings(u.clone(), i.clone()))
    assert abs(float(r.mean())) < 0.02 and abs(float(r.std()) - 0.5) < 0.02
    assert not torch.equal(r, lowrank_ratings(u, i, seed=1))


def test_numa_binding_is_best_effort_and_restricts_to_the_gpu_node(tmp_path):
    import os

    from fps_b200.utils import numa

    assert numa.parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert numa.parse_cpulist("") == set()
    # no GPU / no sysfs entry: nothing happens
    assert numa.bind_to_node(None)["bound"] is False
    assert numa.gpu_numa_node(0, pci_root=str(tmp_path)) is None
    before = os.sched_getaffinity(0)
    cpus = sorted(before)
    if len(cpus) < 4:
        return
    half = len(cpus) // 2
    for n, part in enumerate((cpus[:half], cpus[half:])):
        d = tmp_path / f"node{n}"
        d.mkdir()
        (d / "cpulist").write_text(",".join(str(c) for c in part) + "\n")
    try:
        info = numa.bind_to_node(1, node_root=str(tmp_path), set_policy=False)
        assert info["bound"] and info["cpus"] == len(cpus) - half
        assert os.sched_getaffinity(0) == set(cpus[half:])
        # a node whose CPUs are outside the allowed set leaves the placement alone
        (tmp_path / "node1" / "cpulist").write_text("100000-100003\n")
        info = numa.bind_to_node(1, node_root=str(tmp_path), set_policy=False)
        assert not info["bound"] and os.sched_getaffinity(0) == set(cpus[half:])
    finally:
        os.sched_setaffinity(0, before)
    # single-node box (this container's real sysfs): no-op
    assert numa.bind_to_node(0)["bound"] is False or len(os.listdir("/sys/devices/system/node")) > 1
    os.sched_setaffinity(0, before)
    os.environ["FPS_NUMA_BIND"] = "0"
    try:
        assert numa.bind_to_gpu_node(0)["why"] == "FPS_NUMA_BIND=0"
    finally:
        del os.environ["FPS_NUMA_BIND"]
