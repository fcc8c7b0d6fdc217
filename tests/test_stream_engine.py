"""DataStream routing semantics and engine edge cases."""
import pytest

from fps_b200 import DataStream, ParameterServerLogic, WorkerLogic, transform
from fps_b200.api import CtorFork
from fps_b200.runtime.local_engine import clone_logic
from fps_b200.utils import EOF


class Collect(WorkerLogic):
    seen = None

    def __init__(self):
        self.mine = []

    def onRecv(self, data, ps):
        self.mine.append(data)
        ps.output((id(self), data))

    def onPullRecv(self, *a):
        pass


class NoPS(ParameterServerLogic):
    def onPullRecv(self, *a):
        pass

    def onPushRecv(self, *a):
        pass


def _by_worker(out):
    d = {}
    for wid, x in out.worker_outputs():
        d.setdefault(wid, []).append(x)
    return sorted(d.values(), key=lambda v: (len(v), v))


def test_broadcast_partition_custom_forward_and_operators():
    s = DataStream.from_collection(range(10)).map(lambda x: x * 2).filter(lambda x: x % 4 == 0).flat_map(lambda x: [x, x + 1])
    assert s.collect() == [0, 1, 4, 5, 8, 9, 12, 13, 16, 17]
    out = transform(DataStream.from_collection(range(6)).broadcast(), Collect(), NoPS(), 3, 1, 50)
    assert _by_worker(out) == [[0, 1, 2, 3, 4, 5]] * 3                       # every worker sees every record
    out = transform(DataStream.from_collection(range(9)).partition_custom(lambda k, n: k % n), Collect(), NoPS(), 3, 1, 50)
    assert _by_worker(out) == [[0, 3, 6], [1, 4, 7], [2, 5, 8]]
    out = transform(DataStream.from_parallel([[1, 2], [10], [100, 200, 300]]), Collect(), NoPS(), 3, 1, 50)
    assert _by_worker(out) == [[10], [1, 2], [100, 200, 300]]                # forward: source i -> worker i
    u = DataStream.from_collection([1, 2]).with_eof(EOF).union(DataStream.from_collection([3]))
    out = transform(u, Collect(), NoPS(), 2, 1, 50)
    vals = [x for v in _by_worker(out) for x in v]
    assert vals.count(EOF()) == 2 and sorted(x for x in vals if not isinstance(x, EOF)) == [1, 2, 3]
    with pytest.raises(RuntimeError):
        transform(DataStream.from_collection([1]).partition_custom(lambda k, n: 7), Collect(), NoPS(), 2, 1, 50)


def test_exceptions_in_logic_propagate_and_engine_stops():
    class Boom(WorkerLogic):
        def onRecv(self, data, ps):
            if data == 3:
                raise ValueError("boom")
            ps.pull(data)

        def onPullRecv(self, *a):
            pass

    class PS(ParameterServerLogic):
        def onPullRecv(self, id, w, ps):
            ps.answerPull(id, 0, w)

        def onPushRecv(self, *a):
            pass

    with pytest.raises(ValueError):
        transform(list(range(10)), Boom(), PS(), 2, 2, 50)

    class BoomPS(PS):
        def onPullRecv(self, id, w, ps):
            raise KeyError("server side")

    class Puller(WorkerLogic):
        def onRecv(self, data, ps):
            ps.pull(data)

        def onPullRecv(self, *a):
            pass

    with pytest.raises(KeyError):
        transform(list(range(10)), Puller(), BoomPS(), 2, 2, 50)


def test_per_subtask_copies_and_ctor_fork():
    import threading

    class Stateful(CtorFork, WorkerLogic):
        def __init__(self, tag, scale=1):
            self.tag, self.scale = tag, scale
            self.lock = threading.Lock()          # not deep-copyable: needs fork()
            self.n = 0

        def onRecv(self, data, ps):
            with self.lock:
                self.n += 1
            ps.output((self.tag, self.n))

        def onPullRecv(self, *a):
            pass

    proto = Stateful("x", scale=2)
    c = clone_logic(proto)
    assert c is not proto and (c.tag, c.scale, c.n) == ("x", 2, 0)
    out = transform(list(range(12)), proto, NoPS(), 3, 1, 50)
    counts = sorted(n for _, n in out.worker_outputs())
    assert counts == [1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4] and proto.n == 0      # each subtask has its own state

    class NotCopyable(WorkerLogic):
        def __init__(self):
            self.lock = threading.Lock()

        def onRecv(self, d, ps):
            pass

        def onPullRecv(self, *a):
            pass

    with pytest.raises(TypeError):
        transform([1], NotCopyable(), NoPS(), 2, 1, 50)


def test_stock_protocol_fast_path_equals_the_general_path(monkeypatch):
    """Stock senders / receivers + default partitioners are routed by id with plain queue records (no message
    objects); the observable behaviour is the general path's."""
    from fps_b200 import WorkerLogic, addPullLimiter, transform
    from fps_b200.protocol.senders import SimpleWorkerSender
    from fps_b200.runtime.transform import default_param_partitioner, default_worker_partitioner
    from fps_b200.server.logics import SimplePSLogic

    class W(WorkerLogic):
        def onRecv(self, data, ps):
            ps.pull(data)

        def onPullRecv(self, id, value, ps):
            ps.push(id, 1)
            ps.output((id, value))

    def run(wp, pp, **kw):
        data = [i % 37 for i in range(600)]
        out = transform(data, addPullLimiter(W(), 8), lambda id: 10 * id, lambda p, d: p + d, wp, pp, 50, **kw)
        return out

    monkeypatch.setenv("FPS_ENGINE_FAST", "1")
    a = run(1, 1)
    assert a.engine.used_fast_path
    monkeypatch.setenv("FPS_ENGINE_FAST", "0")
    b = run(1, 1)
    assert not b.engine.used_fast_path
    want = {k: 10 * k + len([i for i in range(600) if i % 37 == k]) for k in range(37)}

    def final(out):
        f = {}
        for k, v in out.ps_outputs():
            f[k] = max(f.get(k, v), v)
        return f

    # (thread interleaving decides the order of the result stream on either path; the content is fixed)
    assert final(a) == final(b) == want
    assert sorted(a.ps_outputs()) == sorted(b.ps_outputs())
    assert len(list(a.worker_outputs())) == len(list(b.worker_outputs())) == 600
    monkeypatch.setenv("FPS_ENGINE_FAST", "1")
    c = run(4, 3)
    assert c.engine.used_fast_path and final(c) == want
    # anything non-stock takes the general path
    class MySender(SimpleWorkerSender):
        pass

    d = transform([1, 2, 3], W(), SimplePSLogic(lambda id: 0, lambda p, q: p + q), default_param_partitioner(2),
                  default_worker_partitioner(2), 2, 2, None, MySender(), None, None, 50)
    assert not d.engine.used_fast_path and len(list(d.ps_outputs())) == 3
    e = transform([1, 2, 3], W(), SimplePSLogic(lambda id: 0, lambda p, q: p + q), lambda m: m.paramId % 2,
                  default_worker_partitioner(2), 2, 2, iterationWaitTime=50)
    assert not e.engine.used_fast_path and len(list(e.ps_outputs())) == 3
