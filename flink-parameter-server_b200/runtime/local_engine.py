"""In-process asynchronous engine: workers and PS shards as threads with FIFO inboxes.

This is the host-tier execution backend ("CPU backend" of SURVEY §7.1-2): it runs arbitrary
Python ``WorkerLogic`` / ``ParameterServerLogic`` callbacks with the reference's semantics --

* the dataflow of ``FlinkParameterServer.transform`` (FPS:340-481): data -> worker ``onRecv``;
  worker requests -> ``paramPartitioner`` -> PS shard -> receiver -> ``onPullRecv/onPushRecv``;
  answers -> ``wInPartition`` (worker index, range checked FPS:455-463) -> worker receiver ->
  ``onPullRecv``; worker outputs ``Left``, PS outputs ``Right``;
* FIFO per (producer, consumer) pair, which the reference's per-id answer queues rely on
  (SURVEY §3.2);
* termination: the reference stops when the feedback edge is idle for ``iterationWaitTime`` ms
  (FPS:49-52, 480).  Here idleness is *detected* (all sources exhausted, no message in flight,
  all batching senders flushed) and then has to persist for ``iterationWaitTime`` ms, so finite
  jobs end promptly while logics with background threads still get their grace period.

It is also the semantic oracle for the device tier's tests.
"""
from __future__ import annotations

import copy
import os
import queue
import threading
import time
from typing import Any, Callable, List, Optional

from ..api import (Left, LooseParameterServerLogic, LooseWorkerLogic, ParameterServer,
                   ParameterServerClient, Right, RuntimeContext)
from ..protocol.messages import PullAnswer
from ..parallel.partitioner import stable_hash
from ..protocol.senders import (PSReceiver, PSSender, SimplePSReceiver, SimplePSSender, SimpleWorkerReceiver,
                                SimpleWorkerSender, WorkerReceiver, WorkerSender)
from .stream import DataStream, ResultStream, as_stream

_DATA, _ANSWER, _REQ, _STOP, _PULL, _PUSH = 0, 1, 2, 3, 4, 5


def assign_subtask(obj: Any, index: int, parallelism: int, depth: int = 0) -> None:
    """Tell a per-subtask logic copy (and the logics it wraps) which subtask it is -- Flink's
    ``getIndexOfThisSubtask``; seeded logics derive their own random stream from it."""
    if obj is None or depth > 4:
        return
    try:
        obj.subtaskIndex, obj.subtaskParallelism = int(index), int(parallelism)
    except Exception:   # objects with __slots__ / read-only attributes
        return
    for name in ("inner", "logic", "workerLogic"):
        sub = obj.__dict__.get(name) if hasattr(obj, "__dict__") else None
        if sub is not None and (hasattr(sub, "onRecv") or hasattr(sub, "onPullRecv")):
            assign_subtask(sub, index, parallelism, depth + 1)


def clone_logic(obj: Any) -> Any:
    """Per-subtask copy of a logic / sender object (Flink serialises one copy per subtask)."""
    if hasattr(obj, "fork") and callable(obj.fork):
        return obj.fork()
    try:
        return copy.deepcopy(obj)
    except Exception as e:
        raise TypeError(
            f"{type(obj).__name__} cannot be copied per subtask ({e}); give it a fork() method "
            "(e.g. inherit fps_b200.api.CtorFork)") from e


class _Activity:
    """In-flight accounting for quiescence detection, without a shared lock on the message path.

    Every thread that sends or finishes messages owns a slot ``[sent, done, last_event]`` that only it
    writes (a single lock here turned into a convoy: all worker / PS threads serialised on it and the
    engine ran at ~13 K messages/s).  The controller sums the ``done`` counters FIRST and the ``sent``
    counters afterwards: ``done == sent`` then implies that at the first instant nothing was in flight
    and nothing was sent in between (every message is sent before it is done)."""

    def __init__(self):
        self.lock = threading.Lock()          # slot registration, sources_open and error only
        self._local = threading.local()
        self._slots: List[list] = []
        self.sources_open = 0
        self.error: Optional[BaseException] = None
        self._t0 = time.monotonic()

    def _slot(self) -> list:
        s = getattr(self._local, "slot", None)
        if s is None:
            s = [0, 0, time.monotonic()]
            self._local.slot = s
            with self.lock:
                self._slots.append(s)
        return s

    def add(self, n: int = 1) -> None:
        try:
            s = self._local.slot
        except AttributeError:
            s = self._slot()
        s[0] += n
        s[2] = time.monotonic()

    def done(self, n: int = 1) -> None:
        try:
            s = self._local.slot
        except AttributeError:
            s = self._slot()
        s[1] += n
        s[2] = time.monotonic()

    def touch(self) -> None:
        self._slot()[2] = time.monotonic()

    @property
    def inflight(self) -> int:
        done = sum(s[1] for s in list(self._slots))
        sent = sum(s[0] for s in list(self._slots))   # later, possibly larger snapshot: never under-counts
        return sent - done

    def idle_for(self) -> float:
        if self.sources_open != 0 or self.inflight != 0:
            return 0.0
        last = max([s[2] for s in list(self._slots)], default=self._t0)
        return time.monotonic() - max(last, self._t0)


class MessagingPSClient(ParameterServerClient):
    """Adapts the client API to the worker sender (FPS:1203-1230); thread safe."""

    def __init__(self, sender: WorkerSender, partitionId: int, emit_to_ps, emit_output):
        self.sender = sender
        self.partitionId = partitionId
        self._emit_to_ps = emit_to_ps
        self._emit_output = emit_output

    def pull(self, id) -> None:
        self.sender.onPull(id, self._emit_to_ps, self.partitionId)

    def push(self, id, deltaUpdate) -> None:
        self.sender.onPush(id, deltaUpdate, self._emit_to_ps, self.partitionId)

    def output(self, out) -> None:
        self._emit_output(Left(out))


class MessagingPS(ParameterServer):
    """Adapts the server API to the PS sender (FPS:1178-1198)."""

    def __init__(self, sender: PSSender, emit_to_worker, emit_output):
        self.sender = sender
        self._emit_to_worker = emit_to_worker
        self._emit_output = emit_output

    def answerPull(self, id, value, workerPartitionIndex) -> None:
        self.sender.onPullAnswer(id, value, workerPartitionIndex, self._emit_to_worker)

    def output(self, out) -> None:
        self._emit_output(Right(out))


class LocalEngine:
    def __init__(self, workerParallelism: int, psParallelism: int, iterationWaitTime: float,
                 call_worker_open: bool = True, ps_batch: int = 4096, fast_path: Optional[bool] = None):
        self.ps_batch = max(1, int(ps_batch))
        # stock protocol (Simple sender / receivers, default partitioners): route by id and queue plain records,
        # no message objects (same observable behaviour; FPS_ENGINE_FAST=0 forces the general path)
        self.fast_path = (os.environ.get("FPS_ENGINE_FAST", "1") != "0") if fast_path is None else bool(fast_path)
        self.wP = int(workerParallelism)
        self.psP = int(psParallelism)
        self.wait_s = max(0.0, float(iterationWaitTime) / 1000.0)
        self.call_worker_open = call_worker_open

    def run(self, trainingData, workerLogic: LooseWorkerLogic, psLogic: LooseParameterServerLogic,
            paramPartitioner: Callable[[Any], int], wInPartition: Callable[[Any], int],
            workerReceiver: WorkerReceiver, workerSender: WorkerSender, psReceiver: PSReceiver,
            psSender: PSSender) -> ResultStream:
        stream: DataStream = as_stream(trainingData)
        act = _Activity()
        results: List[Any] = []
        res_lock = threading.Lock()
        w_inbox = [queue.SimpleQueue() for _ in range(self.wP)]
        ps_inbox = [queue.SimpleQueue() for _ in range(self.psP)]

        def emit_output(x) -> None:
            with res_lock:
                results.append(x)

        def emit_to_ps(msg) -> None:
            dest = int(paramPartitioner(msg)) % self.psP
            act.add()
            ps_inbox[dest].put((_REQ, msg))

        def emit_to_worker(msg) -> None:
            dest = int(wInPartition(msg))
            if not 0 <= dest < self.wP:
                raise RuntimeError("Pull answer key should be the partition ID itself!")
            act.add()
            w_inbox[dest].put((_ANSWER, msg))

        # per-subtask copies, like Flink's serialised operator instances
        w_logic = [clone_logic(workerLogic) for _ in range(self.wP)]
        w_send = [clone_logic(workerSender) for _ in range(self.wP)]
        w_recv = [clone_logic(workerReceiver) for _ in range(self.wP)]
        p_logic = [clone_logic(psLogic) for _ in range(self.psP)]
        for i, lg in enumerate(w_logic):
            assign_subtask(lg, i, self.wP)
        for j, lg in enumerate(p_logic):
            assign_subtask(lg, j, self.psP)
        p_send = [clone_logic(psSender) for _ in range(self.psP)]
        p_recv = [clone_logic(psReceiver) for _ in range(self.psP)]
        for s in w_send:
            if hasattr(s, "bind_partitioner"):
                s.bind_partitioner(lambda m: int(paramPartitioner(m)) % self.psP)
        for s in p_send:
            if hasattr(s, "bind_partitioner"):
                s.bind_partitioner(lambda m: int(wInPartition(m)))
        clients = [MessagingPSClient(w_send[i], i, emit_to_ps, emit_output) for i in range(self.wP)]
        servers = [MessagingPS(p_send[j], emit_to_worker, emit_output) for j in range(self.psP)]
        fast = (self.fast_path and type(workerSender) is SimpleWorkerSender
                and type(workerReceiver) is SimpleWorkerReceiver and type(psSender) is SimplePSSender
                and type(psReceiver) is SimplePSReceiver
                and getattr(paramPartitioner, "fps_default", None) == "hash"
                and getattr(wInPartition, "fps_default", None) == "worker_index")
        self.used_fast_path = fast
        if fast:
            psP, wP, add = self.psP, self.wP, act.add
            ps_put = [q.put for q in ps_inbox]
            w_put = [q.put for q in w_inbox]

            class FastClient(ParameterServerClient):       # MessagingPSClient + SimpleWorkerSender + hash routing
                __slots__ = ("partitionId",)

                def __init__(self, partitionId: int):
                    self.partitionId = partitionId

                def pull(self, id) -> None:
                    dest = stable_hash(id) % psP
                    add()
                    ps_put[dest]((_PULL, id, self.partitionId))

                def push(self, id, deltaUpdate) -> None:
                    dest = stable_hash(id) % psP
                    add()
                    ps_put[dest]((_PUSH, id, deltaUpdate))

                def output(self, out) -> None:
                    emit_output(Left(out))

            class FastServer(ParameterServer):             # MessagingPS + SimplePSSender + worker-index routing
                __slots__ = ()

                def answerPull(self, id, value, workerPartitionIndex) -> None:
                    dest = int(workerPartitionIndex)
                    if not 0 <= dest < wP:
                        raise RuntimeError("Pull answer key should be the partition ID itself!")
                    add()
                    w_put[dest]((_ANSWER, id, value))

                def output(self, out) -> None:
                    emit_output(Right(out))

            clients = [FastClient(i) for i in range(self.wP)]
            servers = [FastServer() for _ in range(self.psP)]

        def fail(e: BaseException) -> None:
            with act.lock:
                if act.error is None:
                    act.error = e

        def worker_loop(i: int) -> None:
            logic, recv, client = w_logic[i], w_recv[i], clients[i]

            def on_answer(a: PullAnswer) -> None:
                logic.onPullRecv(a.paramId, a.param, client)

            try:      # everything that can raise is inside: a dead thread must fail the job, not hang it
                inbox_get, on_answer_msg, done = w_inbox[i].get, recv.onPullAnswerRecv, act.done
                if self.call_worker_open:
                    logic.open()
                on_recv = logic.onRecv          # bound after open(): a logic may rebind its callbacks there
                while True:
                    kind, payload = inbox_get()
                    if kind == _STOP:
                        break
                    try:
                        if kind == _DATA:
                            on_recv(payload, client)
                        else:
                            on_answer_msg(payload, on_answer)
                    finally:
                        done()
            except BaseException as e:  # noqa: BLE001
                fail(e)

        def worker_loop_fast(i: int) -> None:
            logic, client = w_logic[i], clients[i]
            try:
                inbox_get, done = w_inbox[i].get, act.done
                if self.call_worker_open:
                    logic.open()
                on_recv, on_pull_recv = logic.onRecv, logic.onPullRecv
                while True:
                    item = inbox_get()
                    kind = item[0]
                    if kind == _STOP:
                        break
                    try:
                        if kind == _ANSWER:
                            on_pull_recv(item[1], item[2], client)
                        else:
                            on_recv(item[1], client)
                    finally:
                        done()
            except BaseException as e:  # noqa: BLE001
                fail(e)

        def ps_loop_fast(j: int) -> None:
            logic, server = p_logic[j], servers[j]
            inbox = ps_inbox[j]
            try:
                inbox_get, done = inbox.get, act.done
                logic.open({}, RuntimeContext(j, self.psP))
                on_pull, on_push = logic.onPullRecv, logic.onPushRecv
                flush = getattr(logic, "flush", None)
                stop = False
                while not stop:
                    item = inbox_get()
                    if item[0] == _STOP:
                        break
                    n = 1
                    try:
                        if item[0] == _PULL:
                            on_pull(item[1], item[2], server)
                        else:
                            on_push(item[1], item[2], server)
                        if flush is not None:
                            while n < self.ps_batch:
                                try:
                                    item = inbox.get_nowait()
                                except queue.Empty:
                                    break
                                if item[0] == _STOP:
                                    stop = True
                                    break
                                n += 1
                                if item[0] == _PULL:
                                    on_pull(item[1], item[2], server)
                                else:
                                    on_push(item[1], item[2], server)
                            flush(server)
                    finally:
                        done(n)
            except BaseException as e:  # noqa: BLE001
                fail(e)

        def ps_loop(j: int) -> None:
            logic, recv, server = p_logic[j], p_recv[j], servers[j]

            def on_pull(id, widx) -> None:
                logic.onPullRecv(id, widx, server)

            def on_push(id, delta) -> None:
                logic.onPushRecv(id, delta, server)

            # device-resident stores (server/device_logics.py) record requests and execute them in
            # batches: drain whatever has queued up, decode it, then one flush() = a few kernels
            inbox = ps_inbox[j]
            try:
                inbox_get, on_msg, done = ps_inbox[j].get, recv.onWorkerMsg, act.done
                logic.open({}, RuntimeContext(j, self.psP))
                flush = getattr(logic, "flush", None)
                stop = False
                while not stop:
                    kind, payload = inbox_get()
                    if kind == _STOP:
                        break
                    n = 1
                    try:
                        on_msg(payload, on_pull, on_push)
                        if flush is not None:
                            while n < self.ps_batch:
                                try:
                                    kind, payload = inbox.get_nowait()
                                except queue.Empty:
                                    break
                                if kind == _STOP:
                                    stop = True
                                    break
                                n += 1
                                on_msg(payload, on_pull, on_push)
                            flush(server)
                    finally:
                        done(n)
            except BaseException as e:  # noqa: BLE001
                fail(e)

        # ---- sources ---------------------------------------------------------------------
        groups = {}
        for s in stream.sources:
            if s.group is not None:
                groups[s.group] = groups.get(s.group, 0) + 1
        group_lock = threading.Lock()
        act.sources_open = len(stream.sources)

        def feed(src) -> None:
            counter = [src.index]
            try:
                for rec in src.make_iter():
                    if act.error is not None:
                        break
                    for t in src.routing.targets(rec, self.wP, src.index, counter):
                        act.add()
                        w_inbox[t].put((_DATA, rec))
                if src.group is not None:
                    with group_lock:
                        groups[src.group] -= 1
                        last = groups[src.group] == 0
                    if last:
                        marker = stream.eof_markers[src.group]
                        for t in range(self.wP):
                            act.add()
                            w_inbox[t].put((_DATA, marker()))
            except BaseException as e:  # noqa: BLE001
                fail(e)
            finally:
                act.touch()
                with act.lock:
                    act.sources_open -= 1

        threads = [threading.Thread(target=worker_loop_fast if fast else worker_loop, args=(i,), daemon=True,
                                    name=f"fps-worker-{i}") for i in range(self.wP)]
        threads += [threading.Thread(target=ps_loop_fast if fast else ps_loop, args=(j,), daemon=True,
                                     name=f"fps-ps-{j}") for j in range(self.psP)]
        feeders = [threading.Thread(target=feed, args=(s,), daemon=True, name=f"fps-src-{k}")
                   for k, s in enumerate(stream.sources)]
        for t in threads + feeders:
            t.start()

        # ---- termination: detected idleness that persists for iterationWaitTime ----------------
        flushed_idle = False
        while True:
            if act.error is not None:
                break
            idle = act.idle_for()
            if idle > 0.0 or (act.inflight == 0 and act.sources_open == 0):
                if not flushed_idle:
                    # flush batching senders; they may produce new traffic.  Only a FULL pass over every
                    # sender that emitted nothing proves that no message is stranded: a PS thread may
                    # buffer answers to just-flushed pulls right after its sender was visited.
                    emitted = False
                    for i, s in enumerate(w_send):
                        emitted = bool(s.flush(emit_to_ps)) or emitted
                    for j, s in enumerate(p_send):
                        emitted = bool(s.flush(emit_to_worker)) or emitted
                    flushed_idle = not emitted
                    continue
                if idle >= self.wait_s:
                    break
            else:
                flushed_idle = False
            time.sleep(0.0005 if self.wait_s < 0.05 else 0.002)

        for q in w_inbox + ps_inbox:
            q.put((_STOP, None))
        for t in threads:
            t.join(timeout=10.0)
        if act.error is not None:
            for s in w_send + p_send:
                s.close()
            raise act.error
        # close hooks (worker close, then PS close which may dump the model)
        for lg in w_logic:
            lg.close()
        for j, lg in enumerate(p_logic):
            lg.close(servers[j])
        for s in w_send + p_send:
            s.close()
        self.worker_logics, self.ps_logics = w_logic, p_logic
        return ResultStream(results)
