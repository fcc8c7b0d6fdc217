"""End-to-end host-tier jobs on the in-process engine, mirroring the asserting integration tests of
the reference: T/FlinkSimpleStackTest.scala:122-207 (model load), T/FlinkStringIdentifierTest.scala
(string ids / loose types / model load with custom partitioner), T/FlinkCombinationStackTest.scala
(batched senders), T/FlinkParameterServerTest.scala (custom wire format), T/utils/FlinkEOFTest.scala."""
import random
from collections import Counter

import pytest

from fps_b200 import (DataStream, FlinkParameterServer, LooseParameterServerLogic, ParameterServerLogic,
                      WorkerLogic, addPullLimiter, transform, transformLoose, transformWithDoubleModelLoad,
                      transformWithModelLoad)
from fps_b200.api import Left, Right
from fps_b200.protocol import (CombinationPSSender, CombinationWorkerSender, CountLogic, MultiplePSReceiver,
                               MultipleWorkerReceiver, PSReceiver, PSSender, TimerLogic, WorkerReceiver,
                               WorkerSender, all_of, any_of, PullAnswer)
from fps_b200.server import SimplePSLogicWithClose
from fps_b200.utils import EOF, EOFHandler, block, flatMapWithEOF, with_eof

WAIT = 150  # ms


class PullThenPushOne(WorkerLogic):
    def onRecv(self, data, ps):
        ps.pull(data)

    def onPullRecv(self, paramId, paramValue, ps):
        ps.push(paramId, 1)


class DictPS(ParameterServerLogic):
    def __init__(self):
        self.params = {}

    def onPullRecv(self, id, widx, ps):
        ps.answerPull(id, self.params.setdefault(id, 0), widx)

    def onPushRecv(self, id, delta, ps):
        self.params[id] = self.params.setdefault(id, 0) + delta

    def close(self, ps):
        for kv in self.params.items():
            ps.output(kv)


def test_model_load_i_times_10_plus_3():
    num = 50
    init = [(i, i * 10) for i in range(num)]
    data = [i for i in range(num) for _ in range(3)]
    random.Random(0).shuffle(data)
    model_src = DataStream.from_parallel([init[0::2], init[1::2]]).rebalance()
    data_src = DataStream.from_parallel([data[0::3], data[1::3], data[2::3]]).rebalance()
    out = transformWithModelLoad(model_src)(
        data_src, PullThenPushOne(), DictPS(),
        lambda m: m.paramId % 3, lambda m: m.workerPartitionIndex, 4, 3, WAIT)
    assert all(r.is_right for r in out.collect())
    assert sorted(out.ps_outputs()) == [(i, i * 10 + 3) for i in range(num)]


WORDS = "the quick brown fox jumps over the lazy dog the fox".split() * 7


class WordWorker(WorkerLogic):
    def onRecv(self, data, ps):
        ps.push(data, 1)

    def onPullRecv(self, *a):
        raise AssertionError("no pulls expected")


def test_word_count_string_ids_param_init_update():
    out = transform(WORDS, WordWorker(), lambda w: 0, lambda a, b: a + b, 4, 3, WAIT)
    final = {}
    for (w, c) in out.ps_outputs():
        final[w] = max(final.get(w, 0), c)
    assert final == dict(Counter(WORDS))


def test_word_count_loose_types():
    class W(WorkerLogic):
        def onRecv(self, data, ps):
            ps.pull(data)

        def onPullRecv(self, paramId, paramValue, ps):   # PullP = Int
            ps.push(paramId, True)                        # PushP = Boolean

    out = transformLoose(WORDS, W(), lambda w: 0, lambda c, b: c + (1 if b else 0), 3, 2, WAIT)
    final = {}
    for (w, c) in out.ps_outputs():
        final[w] = max(final.get(w, 0), c)
    assert final == dict(Counter(WORDS))


def test_word_count_explicit_loose_logic_with_close():
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

    out = FlinkParameterServer.transform(WORDS, W(), PS(), 4, 4, WAIT)
    assert dict(out.ps_outputs()) == dict(Counter(WORDS))


def test_model_load_string_ids_custom_partitioner():
    init = [(w, 100) for w in set(WORDS)]
    out = transformWithModelLoad(init)(
        WORDS, PullThenPushOne(), DictPS(), lambda m: len(m.paramId) % 2, None, 3, 2, WAIT)
    c = Counter(WORDS)
    assert dict(out.ps_outputs()) == {w: 100 + c[w] for w in c}


def test_double_model_load_routes_worker_entries():
    class W(WorkerLogic):
        def __init__(self):
            self.model = {}

        def updateModel(self, id, param):
            self.model[id] = param

        def onRecv(self, data, ps):
            ps.pull(data)

        def onPullRecv(self, paramId, paramValue, ps):
            ps.output((paramId, paramValue, dict(self.model)))

    model = [Left((i, i * 2)) for i in range(10)] + [Right((100 + w, "w%d" % w)) for w in range(3)]
    out = transformWithDoubleModelLoad(model)(
        DataStream.from_collection(list(range(10))).partition_custom(lambda k, n: k % n),
        W(), DictPS(), None, None, 3, 2, WAIT, workerModelPartitioner=lambda k, n: (k - 100) % n)
    outs = out.worker_outputs()
    assert sorted((i, v) for i, v, _ in outs) == [(i, i * 2) for i in range(10)]
    for i, _, local in outs:                       # worker i%3 holds exactly its own entry
        assert local == {100 + i % 3: "w%d" % (i % 3)}


def test_combination_stack_count_and_timer():
    data = list(range(200))
    out = transform(
        data, PullThenPushOne(), DictPS(),
        lambda m: (m[0] if isinstance(m, list) else m).paramId % 3,
        lambda m: (m[0] if isinstance(m, list) else m).workerPartitionIndex,
        4, 3, MultipleWorkerReceiver(),
        CombinationWorkerSender(all_of, [CountLogic(10), TimerLogic(0.02)]),
        MultiplePSReceiver(),
        CombinationPSSender(any_of, [CountLogic(7), TimerLogic(0.02)]), 300)
    assert dict(out.ps_outputs()) == {i: 1 for i in data}


def test_fully_custom_wire_format():
    """Custom messages: worker->PS ``(is_pull, [widx, id, delta])``, PS->worker ``(widx, [id, value])``."""
    class WS(WorkerSender):
        def onPull(self, id, collect, pid):
            collect((True, [pid, id, 0]))

        def onPush(self, id, d, collect, pid):
            collect((False, [pid, id, d]))

    class PR(PSReceiver):
        def onWorkerMsg(self, msg, onPull, onPush):
            (is_pull, (pid, id, d)) = msg
            onPull(id, pid) if is_pull else onPush(id, d)

    class PSnd(PSSender):
        def onPullAnswer(self, id, value, widx, collect):
            collect((widx, [str(id), str(value)]))

    class WR(WorkerReceiver):
        def onPullAnswerRecv(self, msg, handler):
            handler(PullAnswer(int(msg[1][0]), int(msg[1][1])))

    out = transform(list(range(40)), PullThenPushOne(), DictPS(), lambda m: m[1][1] % 4, lambda m: m[0],
                    4, 4, WR(), WS(), PR(), PSnd(), WAIT)
    assert dict(out.ps_outputs()) == {i: 1 for i in range(40)}


def test_answer_routed_out_of_range_is_an_error():
    with pytest.raises(RuntimeError):
        transform(list(range(5)), PullThenPushOne(), DictPS(), lambda m: 0, lambda m: 99, 2, 1,
                  None, None, None, None, WAIT)


def test_eof_after_all_sources_staggered():
    """T/utils/FlinkEOFTest.scala:15-113 -- 7 sources x 5 records, one consumer, one EOF after all data."""
    import time

    def src(k):
        for j in range(5):
            time.sleep(0.002 * k)
            yield k * 10 + j

    seen = []

    class W(WorkerLogic):
        def onRecv(self, data, ps):
            seen.append(data)

        def onPullRecv(self, *a):
            pass

    s = with_eof(DataStream.from_parallel([src(k) for k in range(7)]).rebalance())
    transform(s, W(), DictPS(), 1, 1, WAIT)
    assert seen.count(EOF()) == 1 and seen[-1] == EOF()
    assert sum(x for x in seen if not isinstance(x, EOF)) == sum(k * 10 + j for k in range(7) for j in range(5))


def test_flat_map_with_eof_handler():
    class F(EOFHandler):
        def __init__(self):
            self.s = 0

        def flatMap(self, v, collect):
            self.s += v

        def onEOF(self, collect):
            collect(("sum", self.s))

    out = flatMapWithEOF(list(range(10)), F(), 2, lambda k, n: k % n, lambda v: v).collect()
    assert sorted(out) == [("sum", 20), ("sum", 25)]


def test_sleep_blocker_delays_first_record():
    import time

    t0 = time.time()
    got = block([1, 2, 3], 120).collect()
    assert got == [1, 2, 3] and 0.9 * 0.12 <= time.time() - t0 <= 1.0


def test_pull_limiter_inside_engine_and_worker_outputs():
    class W(WorkerLogic):
        def onRecv(self, data, ps):
            ps.pull(data)

        def onPullRecv(self, paramId, paramValue, ps):
            ps.output((paramId, paramValue)); ps.push(paramId, 2)

    out = transform(list(range(100)), addPullLimiter(W(), 3), SimplePSLogicWithClose(lambda i: i, lambda a, b: a + b),
                    3, 2, WAIT)
    assert sorted(out.worker_outputs()) == [(i, i) for i in range(100)]
    assert sorted(out.ps_outputs()) == [(i, i + 2) for i in range(100)]
