"""Sender / receiver traits and implementations (FPS:1323-1374; M/client, M/server).

Worker side:  ``WorkerSender.onPull/onPush`` encode requests, ``WorkerReceiver.onPullAnswerRecv``
decodes answers.  Server side: ``PSReceiver.onWorkerMsg`` decodes requests,
``PSSender.onPullAnswer`` encodes answers.  The ``Simple*`` classes send one message per
request; the ``Combination*`` classes batch with count / timer triggers; ``Multiple*`` receivers
unpack batches.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Sequence

from ..api import Left, Right
from .combination import Combinable, CombinationLogic, CountLogic, TimerLogic, any_of
from .messages import PSToWorker, Pull, PullAnswer, Push, WorkerToPS


# ---- traits -----------------------------------------------------------------------------
class WorkerSender:
    def onPull(self, id, collectAnswerMsg: Callable[[Any], None], partitionId: int) -> None:
        raise NotImplementedError

    def onPush(self, id, deltaUpdate, collectAnswerMsg: Callable[[Any], None], partitionId: int) -> None:
        raise NotImplementedError

    def flush(self, collectAnswerMsg) -> None:
        """Emit anything still buffered (no-op for unbatched senders)."""

    def close(self) -> None:
        pass


class WorkerReceiver:
    def onPullAnswerRecv(self, msg, pullHandler: Callable[[PullAnswer], None]) -> None:
        raise NotImplementedError


class PSReceiver:
    def onWorkerMsg(self, msg, onPullRecv: Callable[[Any, int], None],
                    onPushRecv: Callable[[Any, Any], None]) -> None:
        raise NotImplementedError


class PSSender:
    def onPullAnswer(self, id, value, workerPartitionIndex: int, collectAnswerMsg) -> None:
        raise NotImplementedError

    def flush(self, collectAnswerMsg) -> None:
        pass

    def close(self) -> None:
        pass


# ---- simple (unbatched) -----------------------------------------------------------------
class SimpleWorkerSender(WorkerSender):
    """One ``WorkerToPS`` per pull/push (SimpleWorkerSender.scala:6-15)."""

    def onPull(self, id, collectAnswerMsg, partitionId):
        collectAnswerMsg(WorkerToPS(partitionId, Left(Pull(id))))

    def onPush(self, id, deltaUpdate, collectAnswerMsg, partitionId):
        collectAnswerMsg(WorkerToPS(partitionId, Right(Push(id, deltaUpdate))))


class SimpleWorkerReceiver(WorkerReceiver):
    """Unwrap one answer (SimpleWorkerReceiver.scala:6-13)."""

    def onPullAnswerRecv(self, msg: PSToWorker, pullHandler):
        pullHandler(msg.msg)


def _dispatch_worker_msg(m: WorkerToPS, onPullRecv, onPushRecv) -> None:
    body = m.msg
    if isinstance(body, Left) and isinstance(body.value, Pull):
        onPullRecv(body.value.paramId, m.workerPartitionIndex)
    elif isinstance(body, Right) and isinstance(body.value, Push):
        onPushRecv(body.value.paramId, body.value.delta)
    else:  # SimplePSReceiver.scala:17-18
        raise RuntimeError("Parameter server received unknown message.")


class SimplePSReceiver(PSReceiver):
    """Decode one request (SimplePSReceiver.scala:6-22)."""

    def onWorkerMsg(self, msg: WorkerToPS, onPullRecv, onPushRecv):
        _dispatch_worker_msg(msg, onPullRecv, onPushRecv)


class SimplePSSender(PSSender):
    """One ``PSToWorker`` per answer (SimplePSSender.scala:6-12)."""

    def onPullAnswer(self, id, value, workerPartitionIndex, collectAnswerMsg):
        collectAnswerMsg(PSToWorker(workerPartitionIndex, PullAnswer(id, value)))


# ---- batched ----------------------------------------------------------------------------
class MultipleWorkerReceiver(WorkerReceiver):
    """Unpack a batch of answers (MultipleWorkerReceiver.scala:6-13)."""

    def onPullAnswerRecv(self, msg: Sequence[PSToWorker], pullHandler):
        for m in msg:
            pullHandler(m.msg)


class MultiplePSReceiver(PSReceiver):
    """Unpack a batch of requests (MultiplePSReceiver.scala:6-26)."""

    def onWorkerMsg(self, msg: Sequence[WorkerToPS], onPullRecv, onPushRecv):
        for m in msg:
            _dispatch_worker_msg(m, onPullRecv, onPushRecv)


class _PerDestinationBatcher:
    """Shared machinery: one CombinationLogic per destination once a partitioner is bound."""

    def __init__(self, condition, combinables: Sequence[Combinable]):
        self._proto = CombinationLogic(condition, combinables)
        self._logics: Dict[int, CombinationLogic] = {}
        self._partition: Optional[Callable[[Any], int]] = None

    def bind_partitioner(self, fn: Callable[[Any], int]) -> None:
        """Bind ``message -> destination`` so batches become per-destination (K11)."""
        self._partition = fn

    def fork(self):
        """Fresh sender with the same triggers (one per parallel subtask)."""
        new = object.__new__(type(self))
        _PerDestinationBatcher.__init__(new, self._proto.condition,
                                        [c.fork() for c in self._proto.combinables])
        return new

    def _logic_for(self, single_msg) -> CombinationLogic:
        if self._partition is None:
            return self._proto
        dest = self._partition(single_msg)
        lg = self._logics.get(dest)
        if lg is None:
            lg = self._proto.fork()
            self._logics[dest] = lg
        return lg

    def _add(self, single_msg, collect) -> None:
        self._logic_for(single_msg).logic(lambda buf: buf.append(single_msg), collect)

    def flush(self, collectAnswerMsg) -> bool:
        emitted = bool(self._proto.flush(collectAnswerMsg))
        for lg in list(self._logics.values()):
            emitted = bool(lg.flush(collectAnswerMsg)) or emitted
        return emitted

    def close(self) -> None:
        self._proto.close()
        for lg in self._logics.values():
            lg.close()

    # reference-compatible attribute
    @property
    def data(self) -> List[Any]:
        return self._proto.data


class CombinationWorkerSender(_PerDestinationBatcher, WorkerSender):
    """Batched worker sender (CombinationWorkerSender.scala:9-36)."""

    def onPull(self, id, collectAnswerMsg, partitionId):
        self._add(WorkerToPS(partitionId, Left(Pull(id))), collectAnswerMsg)

    def onPush(self, id, deltaUpdate, collectAnswerMsg, partitionId):
        self._add(WorkerToPS(partitionId, Right(Push(id, deltaUpdate))), collectAnswerMsg)


class CombinationPSSender(_PerDestinationBatcher, PSSender):
    """Batched PS sender (CombinationPSSender.scala:9-24)."""

    def onPullAnswer(self, id, value, workerPartitionIndex, collectAnswerMsg):
        self._add(PSToWorker(workerPartitionIndex, PullAnswer(id, value)), collectAnswerMsg)


class CountClientSender(CombinationWorkerSender):
    """Flush after ``max`` requests (CountClientSender.scala:6)."""

    def __init__(self, max: int):
        super().__init__(any_of, [CountLogic(max)])


class TimerClientSender(CombinationWorkerSender):
    """Flush every ``intervalLength`` seconds (TimerClientSender.scala:8)."""

    def __init__(self, intervalLength: float):
        super().__init__(any_of, [TimerLogic(intervalLength)])


class CountPSSender(CombinationPSSender):
    """CountPSSender.scala:6"""

    def __init__(self, max: int):
        super().__init__(any_of, [CountLogic(max)])


class TimerPSSender(CombinationPSSender):
    """TimerPSSender.scala:8"""

    def __init__(self, intervalLength: float):
        super().__init__(any_of, [TimerLogic(intervalLength)])
