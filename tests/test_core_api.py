"""Host-tier contracts, mirroring the assertions of the reference unit tests:
T/WorkerLogicTest.scala, T/server/{SimplePSLogic,LockPSLogicA,LockPSLogicB}Test.scala,
T/SenderReceiverTest.scala."""
import threading
import time

import pytest

from fps_b200 import (ParameterServer, ParameterServerClient, WorkerLogic, WorkerLogicWithFuture,
                      addBlockingPullLimiter, addPullLimiter)
from fps_b200.protocol import (CombinationPSSender, CombinationWorkerSender, CountClientSender,
                               CountLogic, CountPSSender, MultiplePSReceiver, MultipleWorkerReceiver,
                               PSToWorker, Pull, PullAnswer, Push, SimplePSReceiver, SimplePSSender,
                               SimpleWorkerReceiver, SimpleWorkerSender, TimerClientSender, TimerLogic,
                               WorkerToPS, all_of, any_of)
from fps_b200.api import Left, Right, RuntimeContext
from fps_b200.server import (LockPSLogicA, LockPSLogicB, LooseSimplePSLogic,
                             LooseSimplePSLogicWithClose, RangePSLogicWithClose, SimplePSLogic,
                             SimplePSLogicWithClose)


class CountingClient(ParameterServerClient):
    def __init__(self):
        self.pullCounter = 0
        self.pushed = []
        self.outs = []

    def pull(self, id):
        self.pullCounter += 1

    def push(self, id, d):
        self.pushed.append((id, d))

    def output(self, out):
        self.outs.append(out)


class PullOnRecv(WorkerLogic):
    def onRecv(self, data, ps):
        ps.pull(data)

    def onPullRecv(self, paramId, paramValue, ps):
        pass


def test_pull_limiter_limits_pulls():
    """Exact sequence of T/WorkerLogicTest.scala:34-46."""
    ps = CountingClient()
    w = addPullLimiter(PullOnRecv(), 10)
    for x in range(1, 21):
        w.onRecv(x, ps)
    assert ps.pullCounter == 10
    for x in range(1, 6):
        w.onPullRecv(x, -1, ps)
    assert ps.pullCounter == 15
    w.onRecv(21, ps)
    assert ps.pullCounter == 15
    for x in range(6, 22):
        w.onPullRecv(x, -1, ps)
    assert ps.pullCounter == 21


def test_pull_limiter_never_limits_push_and_output():
    class L(WorkerLogic):
        def onRecv(self, data, ps):
            ps.pull(data); ps.push(data, 1); ps.output(data)

        def onPullRecv(self, i, v, ps):
            pass

    ps = CountingClient()
    w = addPullLimiter(L(), 2)
    for x in range(10):
        w.onRecv(x, ps)
    assert ps.pullCounter == 2 and len(ps.pushed) == 10 and len(ps.outs) == 10


def test_blocking_pull_limiter_blocks_until_answer():
    ps = CountingClient()
    w = addBlockingPullLimiter(PullOnRecv(), 3)
    done = threading.Event()

    def producer():
        for x in range(5):
            w.onRecv(x, ps)
        done.set()

    t = threading.Thread(target=producer, daemon=True)
    t.start()
    time.sleep(0.2)
    assert ps.pullCounter == 3 and not done.is_set()
    w.onPullRecv(0, 0, ps); w.onPullRecv(1, 0, ps)
    assert done.wait(2.0)
    assert ps.pullCounter == 5


def test_worker_logic_with_future_forwards_pull_and_completes():
    got = []

    class F(WorkerLogicWithFuture):
        def onDataRecv(self, data, ps):
            ps.pull(data).onComplete(lambda kv: got.append(kv))

    ps = CountingClient()
    w = F()
    w.onRecv(7, ps); w.onRecv(7, ps)
    assert ps.pullCounter == 2
    w.onPullRecv(7, "a", ps); w.onPullRecv(7, "b", ps)
    assert got == [(7, "a"), (7, "b")]


class MockPS(ParameterServer):
    def __init__(self):
        self.answers = []
        self.outs = []

    def answerPull(self, id, value, widx):
        self.answers.append((id, value, widx))

    def output(self, out):
        self.outs.append(out)


def test_simple_ps_logic_init_on_pull_and_emit_on_push():
    """T/server/SimplePSLogicTest.scala:11-39."""
    ps = MockPS()
    lg = SimplePSLogic(lambda i: 23, lambda x, y: y)
    lg.onPullRecv(42, 0, ps)
    assert lg.params[42] == 23 and ps.answers == [(42, 23, 0)]
    lg.onPushRecv(42, 23, ps)
    assert ps.outs == [(42, 23)]
    lg.onPushRecv(7, 5, ps)  # push before pull stores the delta as the value
    assert lg.params[7] == 5


def test_with_close_and_loose_variants():
    ps = MockPS()
    lg = SimplePSLogicWithClose(lambda i: 0, lambda a, b: a + b)
    lg.onPullRecv(1, 0, ps); lg.onPushRecv(1, 2, ps); lg.onPushRecv(1, 3, ps); lg.onPushRecv(9, 4, ps)
    assert ps.outs == []
    lg.close(ps)
    assert sorted(ps.outs) == [(1, 5), (9, 4)]
    ps = MockPS()
    loose = LooseSimplePSLogic(lambda i: 10, lambda a, b: a + (1 if b else 0))
    loose.onPushRecv(3, True, ps)           # first delta for unseen id dropped (parity)
    assert ps.outs == [(3, 10)]
    loose.onPushRecv(3, True, ps)
    assert ps.outs[-1] == (3, 11)
    ps = MockPS()
    lw = LooseSimplePSLogicWithClose(lambda i: 0, lambda a, b: a + len(b), store=lambda b: len(b))
    lw.onPushRecv("k", "abc", ps); lw.onPushRecv("k", "de", ps); lw.close(ps)
    assert ps.outs == [("k", 5)]


def test_range_ps_logic():
    ps = MockPS()
    lg = RangePSLogicWithClose(10, lambda i: i * 100, lambda a, b: a + b)
    lg.open({}, RuntimeContext(2, 3))       # div = 4, last shard holds ids 8, 9
    assert lg.startIndex == 8 and len(lg.params) == 2
    lg.onPullRecv(9, 1, ps)
    assert ps.answers == [(9, 900, 1)]
    lg.onPushRecv(9, 1, ps); lg.onPushRecv(8, 7, ps)
    lg.close(ps)
    assert sorted(ps.outs) == [(8, 7), (9, 901)]


@pytest.mark.parametrize("cls,dup_kept", [(LockPSLogicA, True), (LockPSLogicB, False)])
def test_lock_ps_logic(cls, dup_kept):
    """T/server/LockPSLogic{A,B}Test.scala."""
    ps = MockPS()
    lg = cls(lambda i: 0, lambda a, b: a + b)
    with pytest.raises(RuntimeError):
        lg.onPushRecv(1, 1, ps)                       # push before pull throws
    lg.onPullRecv(1, 0, ps)                           # first pull answered, lock taken
    assert ps.answers == [(1, 0, 0)] and lg.state(1)[0] is True
    lg.onPullRecv(1, 1, ps)                           # second pull queued, not answered
    assert len(ps.answers) == 1 and lg.state(1)[2] == [1]
    lg.onPullRecv(1, 1, ps)                           # duplicate waiter
    assert len(lg.state(1)[2]) == (2 if dup_kept else 1)
    lg.onPushRecv(1, 5, ps)                           # hand to queue head, stay locked
    assert ps.answers[-1] == (1, 5, 1) and lg.state(1)[0] is True and ps.outs == [(1, 5)]
    while lg.state(1)[2]:
        lg.onPushRecv(1, 1, ps)
    lg.onPushRecv(1, 1, ps)                           # empty queue -> unlock
    assert lg.state(1)[0] is False


def test_simple_senders_receivers_pass_through():
    out = []
    SimpleWorkerSender().onPull(3, out.append, 1)
    SimpleWorkerSender().onPush(3, 0.5, out.append, 1)
    assert out == [WorkerToPS(1, Left(Pull(3))), WorkerToPS(1, Right(Push(3, 0.5)))]
    pulls, pushes = [], []
    for m in out:
        SimplePSReceiver().onWorkerMsg(m, lambda i, w: pulls.append((i, w)), lambda i, d: pushes.append((i, d)))
    assert pulls == [(3, 1)] and pushes == [(3, 0.5)]
    ans = []
    SimplePSSender().onPullAnswer(3, 9.0, 1, ans.append)
    assert ans == [PSToWorker(1, PullAnswer(3, 9.0))]
    got = []
    SimpleWorkerReceiver().onPullAnswerRecv(ans[0], got.append)
    assert got == [PullAnswer(3, 9.0)]
    with pytest.raises(RuntimeError):
        SimplePSReceiver().onWorkerMsg(WorkerToPS(0, Left("garbage")), None, None)


def test_count_sender_flushes_at_n():
    batches = []
    s = CountClientSender(5)
    for i in range(12):
        s.onPull(i, batches.append, 0)
    assert [len(b) for b in batches] == [5, 5]
    s.flush(batches.append)
    assert [len(b) for b in batches] == [5, 5, 2]
    pulls = []
    MultiplePSReceiver().onWorkerMsg(batches[0], lambda i, w: pulls.append(i), None)
    assert pulls == [0, 1, 2, 3, 4]
    ps_b = []
    p = CountPSSender(2)
    for i in range(4):
        p.onPullAnswer(i, i * 1.0, 0, ps_b.append)
    got = []
    MultipleWorkerReceiver().onPullAnswerRecv(ps_b[1], got.append)
    assert got == [PullAnswer(2, 2.0), PullAnswer(3, 3.0)]


def test_timer_sender_flushes_within_interval():
    batches = []
    s = TimerClientSender(0.05)
    s.onPull(1, batches.append, 0); s.onPull(2, batches.append, 0)
    assert batches == []
    time.sleep(0.25)
    assert len(batches) == 1 and len(batches[0]) == 2
    s.close()


def test_or_and_combinations_of_count_and_timer():
    b_or, b_and = [], []
    s_or = CombinationWorkerSender(any_of, [CountLogic(3), TimerLogic(0.08)])
    s_and = CombinationWorkerSender(all_of, [CountLogic(3), TimerLogic(0.08)])
    for i in range(3):
        s_or.onPull(i, b_or.append, 0); s_and.onPull(i, b_and.append, 0)
    assert len(b_or) == 1 and b_and == []       # OR fires on count; AND still needs the timer
    time.sleep(0.3)
    assert len(b_and) == 1 and len(b_and[0]) == 3
    s_or.onPull(9, b_or.append, 0)
    time.sleep(0.3)
    assert len(b_or) == 2 and len(b_or[1]) == 1  # OR fires on timer for the straggler
    s_or.close(); s_and.close()


def test_batches_are_per_destination_once_partitioner_bound():
    batches = []
    s = CountClientSender(2)
    s.bind_partitioner(lambda m: m.paramId % 2)
    for i in [0, 1, 2, 3]:
        s.onPull(i, batches.append, 0)
    assert sorted(sorted(m.paramId for m in b) for b in batches) == [[0, 2], [1, 3]]
