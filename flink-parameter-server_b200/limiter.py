"""Worker-side flow control: pull limiters and future-style pulls.

* :func:`addPullLimiter` -- at most ``pullLimit`` unanswered pulls per worker, excess ids wait in a
  FIFO; every answer releases one credit and issues ONE queued pull; pushes/outputs are never
  limited (WL:196-250, loose form WL:274-327).  Exact contract: T/WorkerLogicTest.scala:34-46.
* :func:`addBlockingPullLimiter` -- ``pull`` blocks the caller while the limit is reached; the user
  must pull from a thread other than the one delivering answers (WL:100-174).
* :class:`WorkerLogicWithFuture` -- ``pull`` returns a future completed in ``onPullRecv``
  (WL:342-459).  The reference version never forwards the pull and builds a self-recursive
  callback; this one works.

The device tier expresses the same limiter as a *credit counter on the GPU*: the fused kernels
keep at most ``pullLimit`` row-pulls in flight (``max_inflight_rows``), and the ring path uses
``fps_credit_*`` device counters (ops/csrc/fps_rings.cu).
"""
from __future__ import annotations

import threading
from collections import deque
from concurrent.futures import Future
from typing import Any, Callable, Deque, Dict

from .api import LooseWorkerLogic, ParameterServerClient, WorkerLogic


class _LimitedClient(ParameterServerClient):
    def __init__(self, owner: "_PullLimiter"):
        self._owner = owner
        self.ps: ParameterServerClient = None  # set before every callback (setPS in the reference)

    def pull(self, id) -> None:
        o = self._owner
        with o._lock:
            if o.pullCounter < o.pullLimit:
                o.pullCounter += 1
                self.ps.pull(id)
            else:
                o.pullQueue.append(id)

    def push(self, id, deltaUpdate) -> None:
        self.ps.push(id, deltaUpdate)

    def output(self, out) -> None:
        self.ps.output(out)


class _PullLimiter(LooseWorkerLogic):
    """Strict FIFO pull limiter decorator (shared by the strict and loose entry points)."""

    def __init__(self, workerLogic: LooseWorkerLogic, pullLimit: int):
        self.inner = workerLogic
        self.pullLimit = int(pullLimit)
        self.pullCounter = 0
        self.pullQueue: Deque[Any] = deque()
        self._lock = threading.RLock()
        self.wrappedPS = _LimitedClient(self)

    def open(self) -> None:
        self.inner.open()

    def close(self) -> None:
        self.inner.close()

    def onRecv(self, data, ps) -> None:
        self.wrappedPS.ps = ps
        self.inner.onRecv(data, self.wrappedPS)

    def onPullRecv(self, paramId, paramValue, ps) -> None:
        self.wrappedPS.ps = ps
        self.inner.onPullRecv(paramId, paramValue, self.wrappedPS)
        with self._lock:
            self.pullCounter -= 1
            if self.pullQueue:
                self.wrappedPS.pull(self.pullQueue.popleft())

    # model-load hook used by transformWithDoubleModelLoad (BaseMFWorkerLogic.scala:39-116)
    def updateModel(self, id, param) -> None:
        self.inner.updateModel(id, param)

    def fork(self):
        from .runtime.local_engine import clone_logic

        return _PullLimiter(clone_logic(self.inner), self.pullLimit)

    def __getattr__(self, name):  # expose the wrapped logic's extra attributes
        if name in ("inner", "fork"):
            raise AttributeError(name)
        return getattr(self.inner, name)


def addPullLimiter(workerLogic: LooseWorkerLogic, pullLimit: int) -> LooseWorkerLogic:
    """Decorate ``workerLogic`` so that at most ``pullLimit`` pulls are in flight."""
    return _PullLimiter(workerLogic, pullLimit)


class _BlockingClient(ParameterServerClient):
    def __init__(self, owner: "_BlockingPullLimiter"):
        self._owner = owner
        self.ps: ParameterServerClient = None

    def setPS(self, ps) -> None:
        with self._owner.psLock:
            self.ps = ps

    def pull(self, id) -> None:
        o = self._owner
        with o.canPull:
            while o.pullCounter >= o.pullLimit:
                o.canPull.wait()
            o.pullCounter += 1
            self.ps.pull(id)

    def push(self, id, deltaUpdate) -> None:
        with self._owner.psLock:
            self.ps.push(id, deltaUpdate)

    def output(self, out) -> None:
        with self._owner.psLock:
            self.ps.output(out)


class _BlockingPullLimiter(LooseWorkerLogic):
    def __init__(self, workerLogic: LooseWorkerLogic, pullLimit: int):
        self.inner = workerLogic
        self.pullLimit = int(pullLimit)
        self.pullCounter = 0
        self.psLock = threading.RLock()
        self.canPull = threading.Condition(self.psLock)
        self.wrappedPS = _BlockingClient(self)

    def open(self) -> None:
        self.inner.open()

    def close(self) -> None:
        self.inner.close()

    def onRecv(self, data, ps) -> None:
        self.wrappedPS.setPS(ps)
        self.inner.onRecv(data, self.wrappedPS)

    def onPullRecv(self, paramId, paramValue, ps) -> None:
        self.wrappedPS.setPS(ps)
        self.inner.onPullRecv(paramId, paramValue, self.wrappedPS)
        with self.canPull:
            self.pullCounter -= 1
            self.canPull.notify()

    def updateModel(self, id, param) -> None:
        self.inner.updateModel(id, param)

    def fork(self):
        from .runtime.local_engine import clone_logic

        return _BlockingPullLimiter(clone_logic(self.inner), self.pullLimit)

    def __getattr__(self, name):
        if name in ("inner", "fork"):
            raise AttributeError(name)
        return getattr(self.inner, name)


def addBlockingPullLimiter(workerLogic: LooseWorkerLogic, pullLimit: int) -> LooseWorkerLogic:
    """Like :func:`addPullLimiter` but ``pull`` blocks instead of queueing (WL:100-174)."""
    return _BlockingPullLimiter(workerLogic, pullLimit)


# ------------------------------------------------------------------------------------------
# futures
# ------------------------------------------------------------------------------------------
class PullAnswerFuture(Future):
    """Future of ``(paramId, paramValue)`` (WL:416-459)."""

    def __init__(self, paramId):
        super().__init__()
        self.paramId = paramId

    def pullArrived(self, paramId, param) -> None:
        self.set_result((paramId, param))

    def onComplete(self, f: Callable[[Any], Any]) -> None:
        self.add_done_callback(lambda fut: f(fut.result()))

    @property
    def isCompleted(self) -> bool:
        return self.done()


class PSClientWithFuture:
    def __init__(self, owner: "WorkerLogicWithFuture"):
        self._owner = owner
        self.ps: ParameterServerClient = None

    def pull(self, id) -> PullAnswerFuture:
        fut = PullAnswerFuture(id)
        self._owner._pullWaiter.setdefault(id, deque()).append(fut)
        self.ps.pull(id)  # the reference forgets this call (WL:367-372)
        return fut

    def push(self, id, deltaUpdate) -> None:
        self.ps.push(id, deltaUpdate)

    def output(self, out) -> None:
        self.ps.output(out)


class WorkerLogicWithFuture(WorkerLogic):
    """Worker logic whose pulls return futures; implement :meth:`onDataRecv`."""

    def __init__(self):
        self._pullWaiter: Dict[Any, Deque[PullAnswerFuture]] = {}
        self._psClient = PSClientWithFuture(self)

    def onDataRecv(self, data, ps: PSClientWithFuture) -> None:
        raise NotImplementedError

    def onRecv(self, data, ps) -> None:
        if not hasattr(self, "_pullWaiter"):
            WorkerLogicWithFuture.__init__(self)
        self._psClient.ps = ps
        self.onDataRecv(data, self._psClient)

    def onPullRecv(self, paramId, paramValue, ps) -> None:
        self._psClient.ps = ps
        q = self._pullWaiter[paramId]
        fut = q.popleft()
        if not q:
            del self._pullWaiter[paramId]
        fut.pullArrived(paramId, paramValue)


# snake_case aliases
add_pull_limiter = addPullLimiter
add_blocking_pull_limiter = addBlockingPullLimiter
