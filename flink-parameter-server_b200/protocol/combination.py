"""Message batching engine: count / timer triggers and their boolean combinations.

Capability parity with M/common/{CombinationLogic,Combinable,CountLogic,TimerLogic}.scala.
Differences by design (SURVEY §2.3 note, §7.4):

* buffers are kept **per destination** once a partitioner is bound, so a flushed batch is always
  homogeneous in destination (the reference routes a mixed batch by its first element);
* the timer thread is a daemon, takes the buffer lock, and stops on ``close()`` (the reference's
  never terminates and races with the buffer, TimerLogic.scala:13-26).
"""
from __future__ import annotations

import threading
import time
from typing import Any, Callable, List, Optional, Sequence


class Combinable:
    """A flush trigger (Combinable.scala:5-27)."""

    def __init__(self):
        self._send = False

    def sendCondition(self) -> bool:
        raise NotImplementedError

    def logic(self, adder, callback, collectAnswerMsg) -> None:
        raise NotImplementedError

    def shouldSend(self) -> bool:
        return self._send

    def send(self, callback, collect) -> None:
        self._send = True
        callback(collect)

    def reset(self) -> None:
        self._send = False

    def clear(self) -> None:
        """The buffer was swapped out: forget everything counted / flagged for the old buffer, so the next
        batch is measured from zero (a stale count would make the next flush fire early, a stale
        ``containsData`` would make the timer fire on an empty buffer)."""
        self._send = False

    def fork(self) -> "Combinable":
        raise NotImplementedError

    def close(self) -> None:
        pass


class CountLogic(Combinable):
    """Trigger after ``max`` messages (CountLogic.scala:5-29)."""

    def __init__(self, max: int):
        super().__init__()
        self.max = int(max)
        self.count = 0

    def sendCondition(self) -> bool:
        return self.count >= self.max

    def logic(self, adder, callback, collectAnswerMsg) -> None:
        self.count += 1
        if self.sendCondition():
            self.send(callback, collectAnswerMsg)
            self.count = 0

    def clear(self) -> None:
        super().clear()
        self.count = 0

    def fork(self) -> "CountLogic":
        return CountLogic(self.max)


class TimerLogic(Combinable):
    """Trigger every ``intervalLength`` seconds if data is buffered (TimerLogic.scala:6-51)."""

    def __init__(self, intervalLength: float):
        super().__init__()
        self.interval = float(intervalLength)
        self.containsData = False
        self._thread: Optional[threading.Thread] = None
        self._stop = threading.Event()
        self._owner_lock: Optional[threading.RLock] = None

    def sendCondition(self) -> bool:
        return self.containsData

    def _run(self, callback, collect) -> None:
        while not self._stop.wait(self.interval):
            lock = self._owner_lock
            if lock is not None:
                lock.acquire()
            try:
                if self.sendCondition():
                    self.send(callback, collect)
                    self.containsData = False
            finally:
                if lock is not None:
                    lock.release()

    def logic(self, adder, callback, collectAnswerMsg) -> None:
        self.containsData = True
        if self._thread is None:
            self._thread = threading.Thread(target=self._run, args=(callback, collectAnswerMsg),
                                            daemon=True)
            self._thread.start()

    def clear(self) -> None:
        super().clear()
        self.containsData = False

    def fork(self) -> "TimerLogic":
        return TimerLogic(self.interval)

    def close(self) -> None:
        self._stop.set()


class CombinationLogic:
    """Buffer + triggers + boolean ``condition`` over the triggers (CombinationLogic.scala:6-35)."""

    def __init__(self, condition: Callable[[List[Combinable]], bool], combinables: Sequence[Combinable]):
        self.condition = condition
        self.combinables = list(combinables)
        self.data: List[Any] = []
        self._flushes = 0
        self._lock = threading.RLock()
        for c in self.combinables:
            if isinstance(c, TimerLogic):
                c._owner_lock = self._lock

    def checkAndSend(self, collect: Callable[[List[Any]], None]) -> None:
        with self._lock:
            if self.condition(self.combinables) and self.data:
                batch, self.data = self.data, []
                self._flushes += 1
                for c in self.combinables:
                    c.clear()
                collect(batch)

    def logic(self, func: Callable[[List[Any]], None], collect: Callable[[List[Any]], None]) -> None:
        with self._lock:
            func(self.data)
            before = self._flushes
            for c in self.combinables:
                c.logic(func, self.checkAndSend, collect)
            if self._flushes != before and not self.data:
                # a trigger fired in the middle of the pass: the triggers visited after it have just
                # accounted for a message that is already gone
                for c in self.combinables:
                    c.clear()

    def flush(self, collect: Callable[[List[Any]], None]) -> bool:
        """Unconditional flush (used at termination so no message is stranded).  Returns whether
        anything was emitted."""
        with self._lock:
            if self.data:
                batch, self.data = self.data, []
                for c in self.combinables:
                    c.clear()
                collect(batch)
                return True
        return False

    def fork(self) -> "CombinationLogic":
        return CombinationLogic(self.condition, [c.fork() for c in self.combinables])

    def close(self) -> None:
        for c in self.combinables:
            c.close()


def any_of(combinables: List[Combinable]) -> bool:
    """OR combination: flush when any trigger fired."""
    return any(c.shouldSend() for c in combinables)


def all_of(combinables: List[Combinable]) -> bool:
    """AND combination: flush only when every trigger fired."""
    return all(c.shouldSend() for c in combinables)
