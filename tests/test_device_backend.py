"""``transform(..., backend="device")``: the reference's asserting stack tests on device-resident stores
(T/FlinkSimpleStackTest.scala:122-207 model load ``i*10+3`` with wP=4 / psP=3;
T/FlinkStringIdentifierTest.scala:37-270 string ids, loose types, custom partitioner), the lock stores
(T/server/LockPSLogicATest.scala, LockPSLogicBTest.scala) and the tensor tier's front door."""
import operator
import random
import warnings
from collections import Counter

import numpy as np
import pytest
import torch

from fps_b200 import (DataStream, LooseParameterServerLogic, WorkerLogic, transform, transformLoose,
                      transformWithModelLoad)
from fps_b200.api import BatchedWorkerLogic
from fps_b200.server import (LockPSLogicA, LockPSLogicB, RangePSLogicWithClose, SimplePSLogic,
                             SimplePSLogicWithClose)
from fps_b200.server.device_logics import (OP_ADD, OP_ASSIGN, OP_HOST, OP_MAX, ValueCodec, classify_update,
                                           to_device_logic)

WAIT = 150
WORDS = "the quick brown fox jumps over the lazy dog the fox".split() * 7


# ---- host-only pieces (run on CPU) -------------------------------------------------------------------
def test_update_classification_and_codec():
    c = ValueCodec(); c.learn(0)
    assert classify_update("add", c) == OP_ADD and classify_update(operator.add, c) == OP_ADD
    assert classify_update(lambda a, b: a + b, c, 0, 1) == OP_ADD
    assert classify_update(lambda cnt, flag: cnt + (1 if flag else 0), c, 0, True) == OP_ADD
    assert classify_update(lambda cnt, flag: cnt + (1 if flag else 0), c, 0, False) == OP_ADD
    assert classify_update(lambda a, b: b, c, 0, 5) == OP_ASSIGN
    assert classify_update(lambda a, b: max(a, b), c, 0, 5) == OP_MAX
    assert classify_update(lambda a, b: a * 2 + b, c, 0, 5) == OP_HOST
    v = ValueCodec(); v.learn([0.0, 0.0, 0.0])
    assert classify_update(lambda a, b: [x + y for x, y in zip(a, b)], v, [0.0] * 3, [1.0, 2.0, 3.0]) == OP_ADD
    assert c.decode(np.array([41.6], dtype=np.float32)) == 42 and v.decode(np.arange(3.0)) == [0.0, 1.0, 2.0]
    assert to_device_logic(SimplePSLogic(lambda i: 0, operator.add)).kind == "simple"
    assert to_device_logic(object()) is None


def test_device_store_engine_path_on_the_cpu_emulation(monkeypatch):
    """The same jobs with the five device operations emulated by torch index ops on host tensors
    (``FPS_DEVICE_STORE_EMULATE=1``): the batching / ordering / store logic is covered by the CPU suite."""
    monkeypatch.setenv("FPS_DEVICE_STORE_EMULATE", "1")
    for job in (test_model_load_i_times_10_plus_3_on_device_stores,
                test_word_count_string_ids_param_init_update_on_device, test_word_count_loose_types_on_device,
                test_model_load_string_ids_custom_partitioner_on_device,
                test_vector_params_and_host_fallback_update_match_the_host_engine, test_range_store_on_device):
        job()
    for cls, dup in ((LockPSLogicA, True), (LockPSLogicB, False)):
        test_lock_stores_on_device(cls, dup)


# ---- device --------------------------------------------------------------------------------------------
gpu = pytest.mark.gpu


class PullThenPushOne(WorkerLogic):
    def onRecv(self, data, ps):
        ps.pull(data)

    def onPullRecv(self, paramId, paramValue, ps):
        ps.push(paramId, 1)


@gpu
def test_model_load_i_times_10_plus_3_on_device_stores():
    num = 50
    init = [(i, i * 10) for i in range(num)]
    data = [i for i in range(num) for _ in range(3)]
    random.Random(0).shuffle(data)
    model_src = DataStream.from_parallel([init[0::2], init[1::2]]).rebalance()
    data_src = DataStream.from_parallel([data[0::3], data[1::3], data[2::3]]).rebalance()
    out = transformWithModelLoad(model_src)(
        data_src, PullThenPushOne(), SimplePSLogicWithClose(lambda i: 0, lambda a, b: a + b),
        lambda m: m.paramId % 3, lambda m: m.workerPartitionIndex, 4, 3, WAIT, backend="device")
    assert sorted(out.ps_outputs()) == [(i, i * 10 + 3) for i in range(num)]
    stores = [lg.inner for lg in out.engine.ps_logics]
    assert all(s.is_device_store and s.stats["host_updates"] == 0 for s in stores)     # REDG path, not host RMW
    assert sum(s.stats["pushes"] for s in stores) >= 3 * num
    assert all(s.rows.is_cuda or s.emulate for s in stores)


class WordPuller(WorkerLogic):
    def onRecv(self, data, ps):
        ps.pull(data)

    def onPullRecv(self, paramId, paramValue, ps):
        ps.push(paramId, 1)
        ps.output(("???", 0))


def _max_per_word(out):
    final = {}
    for x in out.collect():
        w, c = x.value
        final[w] = max(final.get(w, 0), c)
    return final


@gpu
def test_word_count_string_ids_param_init_update_on_device():
    out = transform(WORDS, WordPuller(), lambda w: 0, lambda a, b: a + b, 4, 4, WAIT, backend="device")
    expected = dict(Counter(WORDS)); expected["???"] = 0
    assert _max_per_word(out) == expected
    assert all(s.stats["host_updates"] == 0 for s in out.engine.ps_logics)


@gpu
def test_word_count_loose_types_on_device():
    class W(WorkerLogic):
        def onRecv(self, data, ps):
            ps.pull(data)

        def onPullRecv(self, paramId, paramValue, ps):   # PullP = Int
            assert isinstance(paramValue, int)
            ps.push(paramId, True)                        # PushP = Boolean

    out = transformLoose(WORDS, W(), lambda w: 0, lambda c, b: c + 1 if b else c, 3, 2, WAIT, backend="device")
    assert _max_per_word(out) == dict(Counter(WORDS))


@gpu
def test_user_defined_server_logic_stays_on_host_with_a_warning():
    class PS(LooseParameterServerLogic):
        def __init__(self):
            self.c = {}

        def onPullRecv(self, id, widx, ps):
            ps.answerPull(id, self.c.get(id, 0), widx)

        def onPushRecv(self, id, delta, ps):
            self.c[id] = self.c.get(id, 0) + (1 if delta else 0)

        def close(self, ps):
            for kv in self.c.items():
                ps.output(kv)

    class W(WorkerLogic):
        def onRecv(self, data, ps):
            ps.push(data, True)

        def onPullRecv(self, *a):
            pass

    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        out = transform(WORDS, W(), PS(), 4, 4, WAIT, backend="device")
    assert any("host tier" in str(x.message) for x in w)
    assert dict(out.ps_outputs()) == dict(Counter(WORDS))


@gpu
def test_model_load_string_ids_custom_partitioner_on_device():
    init = [(w, 100) for w in set(WORDS)]
    out = transformWithModelLoad(init)(
        WORDS, PullThenPushOne(), SimplePSLogicWithClose(lambda w: 0, operator.add),
        lambda m: len(m.paramId) % 2, None, 3, 2, WAIT, backend="device")
    c = Counter(WORDS)
    assert dict(out.ps_outputs()) == {w: 100 + c[w] for w in c}
    stores = [lg.inner for lg in out.engine.ps_logics]
    for j, s in enumerate(stores):                      # every word lives on the shard the user function chose
        assert all(len(w) % 2 == j for w in s.keys)


@gpu
def test_vector_params_and_host_fallback_update_match_the_host_engine():
    data = [i % 7 for i in range(60)]

    class W(WorkerLogic):
        def onRecv(self, d, ps):
            ps.pull(d)

        def onPullRecv(self, pid, val, ps):
            ps.push(pid, [0.5, float(pid)])

    init = lambda i: [float(i), 1.0]
    for upd in (lambda a, b: [x + y for x, y in zip(a, b)], lambda a, b: [0.5 * x + y for x, y in zip(a, b)]):
        host = transform(data, W(), SimplePSLogicWithClose(init, upd), 1, 1, WAIT)
        dev = transform(data, W(), SimplePSLogicWithClose(init, upd), 1, 1, WAIT, backend="device")
        h, d = dict(host.ps_outputs()), dict(dev.ps_outputs())
        assert h.keys() == d.keys()
        for k in h:
            np.testing.assert_allclose(d[k], h[k], rtol=1e-5)
    assert dev.engine.ps_logics[0].stats["host_updates"] > 0       # the second update is not a registered op


@gpu
def test_range_store_on_device():
    data = list(range(0, 30, 3)) * 2
    out = transform(data, PullThenPushOne(), RangePSLogicWithClose(30, lambda i: i, operator.add),
                    lambda m: min(m.paramId // 10, 2), lambda m: m.workerPartitionIndex, 2, 3,
                    iterationWaitTime=WAIT, backend="device")
    assert sorted(out.ps_outputs()) == [(i, i + 2) for i in range(0, 30, 3)]


class _MockPS:
    def __init__(self):
        self.answers, self.outputs = [], []

    def answerPull(self, id, value, widx):
        self.answers.append((id, value, widx))

    def output(self, o):
        self.outputs.append(o)


@gpu
@pytest.mark.parametrize("cls,dup_kept", [(LockPSLogicA, True), (LockPSLogicB, False)])
def test_lock_stores_on_device(cls, dup_kept):
    from fps_b200.api import RuntimeContext

    lg = to_device_logic(cls(lambda i: 23, lambda a, b: a + b))
    lg.open({}, RuntimeContext(0, 1))
    ps = _MockPS()
    lg.onPushRecv(42, 1, ps)
    with pytest.raises(RuntimeError):
        lg.flush(ps)                                          # push before pull
    lg.onPullRecv(42, 0, ps); lg.flush(ps)
    assert ps.answers == [(42, 23, 0)] and lg.state(42) == (True, 23, [])
    lg.onPullRecv(42, 1, ps); lg.onPullRecv(42, 1, ps); lg.flush(ps)
    assert len(ps.answers) == 1                               # queued, not answered
    assert lg.state(42)[2] == ([1, 1] if dup_kept else [1])
    lg.onPushRecv(42, 5, ps); lg.flush(ps)                    # hands the fresh value to the queue head
    assert ps.answers[-1] == (42, 28, 1) and lg.state(42)[0] is True and ps.outputs[-1] == (42, 28)
    while lg.state(42)[2]:
        lg.onPushRecv(42, 1, ps); lg.flush(ps)
    lg.onPushRecv(42, 1, ps); lg.flush(ps)
    assert lg.state(42)[0] is False                           # empty queue: unlocked


class PullPushOnceBatched(BatchedWorkerLogic):
    def onRecvBatch(self, batch, ps):
        ps.pull(batch)

    def onPullRecvBatch(self, ids, values, ps):
        ps.push(ids, torch.ones(ids.numel(), values.shape[1], device=ids.device))
        ps.output(int(ids.numel()))


@gpu
@pytest.mark.parametrize("custom", [False, True])
def test_tensor_tier_front_door_model_load_wp4_psp3(custom):
    dev = torch.device("cuda", 0)
    num = 50
    model = [(i, [i * 10.0, -1.0]) for i in range(num)]
    g = torch.Generator().manual_seed(3)
    batches = [torch.randperm(num, generator=g).to(dev) for _ in range(3)]     # every id pulled + pushed 3x
    out = transformWithModelLoad(model)(
        batches, PullPushOnceBatched(), ("zeros", "add", dict(num_ids=num, dim=2, pull_limit=16)),
        (lambda i: (i * 7 + 1) % 3) if custom else None, None, 4, 3, backend="device")
    got = {i: v.tolist() for i, v in out.ps_outputs()}
    assert got == {i: [i * 10.0 + 3, 2.0] for i in range(num)}
    assert sum(out.worker_outputs()) == 3 * num
    t = out.table
    assert t.n_shards == 3 and (t.mode == 2) == custom
    if custom:                                                 # rows live where the user function put them
        ids, _ = t.dump_local(only_touched=False)
        for s in range(3):
            assert all((int(i) * 7 + 1) % 3 == s for i in t.shard_ids(s).tolist())
    t.close()


@gpu
def test_device_credit_counter_pull_limiter():
    from fps_b200.store.sharded_table import ShardedTable

    dev = torch.device("cuda", 0)
    t = ShardedTable(20000, 64, seed=4, init_range=(-1, 1))
    ids = torch.randint(0, 20000, (300000,), device=dev)
    free = t.pull(ids)
    limited = t.pull(ids, pull_limit=64)                      # at most 64 un-answered row pulls at any time
    torch.cuda.synchronize()
    assert torch.equal(free, limited)
    c = t._credits(64, dev)
    assert int(c[0]) == 64 and int(c[1]) >= 0                 # every credit returned (c[1] counts the stalls)
    tiny = t.pull(ids[:50000], pull_limit=2)                  # two credits: the counter must bind, and free them
    torch.cuda.synchronize()
    assert torch.equal(tiny, free[:50000])
    c2 = t._credits(2, dev)
    assert int(c2[0]) == 2 and int(c2[1]) > 0
    t.close()


@gpu
def test_fused_mf_kernel_consumes_the_device_credit_counter():
    """pullLimit inside the fused pull+SGD+push kernel is a device credit counter (WL:196-250), not launch
    geometry: same result as the unlimited kernel on a conflict-free batch, every credit returned."""
    from fps_b200.models.mf.device import DeviceOnlineMF

    dev = torch.device("cuda", 0)
    nu, ni, k, b = 6000, 5000, 64, 3000
    free = DeviceOnlineMF(nu, ni, k, range_min=-0.5, range_max=0.5, learning_rate=0.05, seed=5)
    lim = DeviceOnlineMF(nu, ni, k, range_min=-0.5, range_max=0.5, learning_rate=0.05, seed=5, pull_limit=64)
    g = torch.Generator().manual_seed(1)
    users = torch.randperm(nu, generator=g)[:b].int().to(dev)
    items = torch.randperm(ni, generator=g)[:b].int().to(dev)
    ratings = torch.rand(b, generator=g).to(dev)
    free.step(users, items, ratings); lim.step(users, items, ratings)
    torch.cuda.synchronize()
    torch.testing.assert_close(lim.users, free.users, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(lim.items.local, free.items.local, rtol=1e-6, atol=1e-7)
    assert lim.credits.tolist()[0] == 64 and lim.credits.tolist()[1] >= 0     # all returned; [1] = stalls
    assert lim.stats[1].item() == b
    free.close(); lim.close()
