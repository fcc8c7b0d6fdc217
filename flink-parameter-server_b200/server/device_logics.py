"""Device-resident versions of the built-in server stores, behind the per-record PS API.

``transform(..., backend="device")`` replaces the host ``ParameterServerLogic`` of every PS shard by a
:class:`DeviceStoreLogic`: the shard's parameters live in a growable ``[capacity, stride]`` fp32 block
in the HBM of GPU ``shard % n_gpus``; requests that arrive at the shard (already routed by the job's
``paramPartitioner`` -- hash, range or an arbitrary user function -- exactly like FPS:416-420) are
*recorded* by ``onPullRecv`` / ``onPushRecv`` and executed in batches by ``flush()``:

* the shard's interner maps any hashable id (ints, strings, tuples) to a row slot -- the device address
  of a parameter is ``(shard, slot)``, i.e. the ``(owner, slot)`` lookup that realises custom
  partitioners and opaque ids on dense device tables;
* a run of pulls is ONE gather kernel (``fps_pull_gather``), a run of pushes is ONE ``red.add`` /
  fetch-add / assign kernel (``fps_push_add``, ``fps_push_add_fetch``, ``fps_push_assign``) on the
  shard's CUDA stream; answers and PS outputs are decoded from one D2H copy per run;
* request order is preserved: runs execute in arrival order, and a run is split where a
  non-commutative update would see the same id twice.

Store semantics reproduced (M/server/*.scala): lazy ``paramInit`` on first pull; push to an unseen id
stores the delta (Simple), ``init(id)`` (LooseSimple, first delta dropped) or ``store(delta)``
(LooseSimpleWithClose); ``(id, newValue)`` PS output on every push (Simple / LooseSimple) or the model
dump at ``close()`` (``*WithClose``, Range); per-key locks with waiter queues (Lock A / B).

``paramUpdate`` is executed on the device when it is one of the registered ops -- ``"add"`` (also
``operator.add``, ``vectorSum``, and any pure function that probes as ``x + d``), ``"assign"``, ``"max"``,
``"min"`` -- and otherwise falls back to a host read-modify-write of the device row (correct for any
pure Python function, slower).
"""
from __future__ import annotations

import math
import operator
from collections import deque
from typing import Any, Callable, Dict, List, Optional, Tuple

import numpy as np

from ..api import LooseParameterServerLogic, RuntimeContext

OP_ADD, OP_ASSIGN, OP_MAX, OP_MIN, OP_HOST = "add", "assign", "max", "min", "host"
_SYMBOLIC = {"add": OP_ADD, "sum": OP_ADD, "assign": OP_ASSIGN, "set": OP_ASSIGN, "max": OP_MAX, "min": OP_MIN}


# ------------------------------------------------------------------------------------------------
# value codec: parameter / delta values  <->  fp32 rows
# ------------------------------------------------------------------------------------------------
class ValueCodec:
    """Encodes scalars (bool / int / float) and vectors (list / tuple / numpy / torch) as fp32 rows of a
    fixed ``dim`` and decodes rows back into the kind of value first seen for pulls."""

    def __init__(self):
        self.dim: Optional[int] = None
        self.kind: Optional[str] = None   # 'int' | 'float' | 'list' | 'tuple' | 'ndarray' | 'tensor'

    @staticmethod
    def _flat(v) -> np.ndarray:
        if hasattr(v, "detach"):
            v = v.detach().cpu().numpy()
        return np.asarray(v, dtype=np.float32).reshape(-1)

    def learn(self, v) -> None:
        if self.kind is None:
            if isinstance(v, bool) or isinstance(v, (int, np.integer)):
                self.kind = "int"
            elif isinstance(v, (float, np.floating)):
                self.kind = "float"
            elif hasattr(v, "detach"):
                self.kind = "tensor"
            elif isinstance(v, np.ndarray):
                self.kind = "ndarray"
            elif isinstance(v, tuple):
                self.kind = "tuple"
            else:
                self.kind = "list"
        if self.dim is None:
            self.dim = int(self._flat(v).size)

    def encode(self, v) -> np.ndarray:
        a = self._flat(v)
        if self.dim is None:
            self.dim = int(a.size)
        if a.size != self.dim:
            raise ValueError(f"parameter values must have {self.dim} elements, got {a.size}")
        return a

    def decode(self, row: np.ndarray):
        k = self.kind
        if k == "int":
            return int(round(float(row[0])))
        if k == "float":
            return float(row[0])
        if k == "tensor":
            import torch

            return torch.from_numpy(np.array(row[: self.dim], dtype=np.float32))
        if k == "ndarray":
            return np.array(row[: self.dim], dtype=np.float64)
        if k == "tuple":
            return tuple(float(x) for x in row[: self.dim])
        return [float(x) for x in row[: self.dim]]


def classify_update(fn, codec: ValueCodec, sample_value=None, sample_delta=None) -> str:
    """Map ``paramUpdate`` onto a registered device op, or ``OP_HOST``.

    Identity first (strings, ``operator.add``, ``vectorSum``, numpy / torch add); otherwise the
    function is *probed* on a few random points of the value kind (it must be pure, as in the
    reference where it is a serialised closure): ``fn(x, d) == x + d`` -> add, ``== d`` -> assign,
    ``== max / min(x, d)``; anything else runs on the host."""
    if isinstance(fn, str):
        if fn not in _SYMBOLIC:
            raise ValueError(f"unknown paramUpdate op {fn!r}; registered: {sorted(_SYMBOLIC)}")
        return _SYMBOLIC[fn]
    if fn is operator.add or fn is np.add or getattr(fn, "__name__", "") == "vectorSum":
        return OP_ADD
    try:
        import torch

        if fn is torch.add:
            return OP_ADD
    except Exception:  # pragma: no cover
        pass
    if sample_value is None or sample_delta is None:
        return OP_HOST
    rng = np.random.RandomState(12345)
    scalar_delta = isinstance(sample_delta, (bool, int, float, np.integer, np.floating))

    def like(sample, arr):
        if isinstance(sample, bool):
            return bool(sample)
        if isinstance(sample, (int, np.integer)):
            return int(round(float(arr[0])))
        if isinstance(sample, (float, np.floating)):
            return float(arr[0])
        if hasattr(sample, "detach"):
            import torch

            return torch.from_numpy(arr.astype(np.float32))
        if isinstance(sample, np.ndarray):
            return arr.astype(sample.dtype)
        return type(sample)(float(x) for x in arr)

    verdicts = []
    n = max(1, codec.dim or 1)
    for t in range(4):
        xa = np.round(rng.uniform(-8, 8, n)) if codec.kind == "int" else rng.uniform(-2, 2, n)
        da = (ValueCodec._flat(sample_delta) if (scalar_delta or t == 0)
              else np.round(rng.uniform(-8, 8, n)) if isinstance(np.asarray(sample_delta).flat[0], (int, np.integer))
              else rng.uniform(-2, 2, n))
        x, d = like(sample_value, xa), like(sample_delta, da)
        try:
            y = ValueCodec._flat(fn(x, d))
        except Exception:
            return OP_HOST
        xe, de = ValueCodec._flat(x), ValueCodec._flat(d)
        if y.shape != xe.shape:
            return OP_HOST
        v = set()
        if np.allclose(y, xe + de, rtol=1e-6, atol=1e-6):
            v.add(OP_ADD)
        if np.allclose(y, de, rtol=1e-6, atol=1e-6):
            v.add(OP_ASSIGN)
        if np.allclose(y, np.maximum(xe, de), rtol=1e-6, atol=1e-6):
            v.add(OP_MAX)
        if np.allclose(y, np.minimum(xe, de), rtol=1e-6, atol=1e-6):
            v.add(OP_MIN)
        verdicts.append(v)
    common = set.intersection(*verdicts) if verdicts else set()
    for op in (OP_ADD, OP_ASSIGN, OP_MAX, OP_MIN):
        if op in common:
            return op
    return OP_HOST


# ------------------------------------------------------------------------------------------------
# the device shard
# ------------------------------------------------------------------------------------------------
class DeviceStoreLogic(LooseParameterServerLogic):
    """One PS shard on a GPU.  ``kind``: ``simple`` | ``simple_close`` | ``loose`` | ``loose_close`` |
    ``range_close`` | ``lockA`` | ``lockB`` (the stores of M/server/*.scala)."""

    is_device_store = True

    def __init__(self, kind: str, paramInit: Callable[[Any], Any], paramUpdate, *,
                 store: Optional[Callable[[Any], Any]] = None, drop_first_delta: bool = True,
                 featureCount: Optional[int] = None, capacity: int = 1024, max_run: int = 1 << 16):
        self.kind = kind
        self.init, self.update, self.store = paramInit, paramUpdate, store
        self.drop_first_delta = drop_first_delta
        self.featureCount = featureCount
        self.capacity0, self.max_run = int(capacity), int(max_run)
        self.codec = ValueCodec()
        self.op: Optional[str] = None
        self._op_cache: Dict[Any, str] = {}
        self.slots: Dict[Any, int] = {}
        self.keys: List[Any] = []
        self.rows = None                # torch [capacity, stride] on the shard's GPU
        self.table_c = None
        self.ops: List[Tuple[int, Any, Any]] = []   # (0 pull | 1 push, id, widx | delta)
        self.locks: Dict[Any, Tuple[bool, deque]] = {}
        self.stats = {"pull_runs": 0, "push_runs": 0, "pulls": 0, "pushes": 0, "host_updates": 0,
                      "kernels": 0}
        self.device = None
        self.stream = None

    def fork(self):
        return DeviceStoreLogic(self.kind, self.init, self.update, store=self.store,
                                drop_first_delta=self.drop_first_delta, featureCount=self.featureCount,
                                capacity=self.capacity0, max_run=self.max_run)

    # -- lifecycle ------------------------------------------------------------------------------
    def open(self, parameters, runtimeContext: RuntimeContext):
        import os
        import torch

        idx = runtimeContext.getIndexOfThisSubtask()
        self.shard, self.n_shards = idx, runtimeContext.getNumberOfParallelSubtasks()
        # FPS_DEVICE_STORE_EMULATE=1: the same store logic over host tensors and torch index ops -- the
        # "fake backend" used by the CPU test-suite to exercise the engine path without a GPU
        self.emulate = os.environ.get("FPS_DEVICE_STORE_EMULATE", "0") == "1"
        if self.emulate:
            self.device, self.stream = torch.device("cpu"), None
        else:
            if not torch.cuda.is_available():
                raise RuntimeError('backend="device" needs a CUDA device; use backend="local" on CPU')
            self.device = torch.device("cuda", idx % torch.cuda.device_count())
            self.stream = torch.cuda.Stream(device=self.device)
        if self.kind == "range_close":
            n, fc = self.n_shards, int(self.featureCount)
            div = int(math.ceil(fc / n))
            mod = fc - (n - 1) * div
            self.range_size = max(mod if (mod != 0 and idx + 1 == n) else div, 0)
            self.range_start = idx * div

    # ---- the five device operations (kernels on CUDA, torch index ops when emulating) -------------------
    def _ctx(self):
        import contextlib
        import torch

        if self.emulate:
            return contextlib.nullcontext()
        stack = contextlib.ExitStack()
        stack.enter_context(torch.cuda.device(self.device))
        stack.enter_context(torch.cuda.stream(self.stream))
        return stack

    def _k_gather(self, ids, out):
        if self.emulate:
            out.copy_(self.rows[ids, : out.shape[1]])
        else:
            from ..ops import native
            native.pull_gather(self.table_c, ids, out)

    def _k_assign(self, ids, vals):
        if self.emulate:
            self.rows[ids, : vals.shape[1]] = vals
        else:
            from ..ops import native
            native.push_assign(self.table_c, ids, vals)

    def _k_add(self, ids, d, fetch: bool):
        if self.emulate:
            if not fetch:
                self.rows.index_add_(0, ids, torch_pad(d, self.rows.shape[1]))
                return None
            out = []
            for i, row in zip(ids.tolist(), d):      # returning adds: every push sees its own prefix sum
                self.rows[i, : d.shape[1]] += row
                out.append(self.rows[i, : self.codec.dim].clone())
            import torch
            return torch.stack(out)
        from ..ops import native
        if fetch:
            return native.push_add_fetch(self.table_c, ids, d)
        native.push_add(self.table_c, ids, d)
        return None

    def _alloc(self, min_rows: int) -> None:
        import torch
        from ..ops import native

        dim = self.codec.dim
        stride = (dim + 3) // 4 * 4
        cap = self.capacity0 if self.rows is None else self.rows.shape[0]
        while cap < min_rows:
            cap *= 2
        if self.rows is not None and cap == self.rows.shape[0]:
            return
        with self._ctx():
            new = torch.zeros((cap, stride), dtype=torch.float32, device=self.device)
            if self.rows is not None:
                new[: self.rows.shape[0]] = self.rows
            self.rows = new
            self.table_c = None if self.emulate else native.local_table(self.rows, dim)

    def _slot_of(self, id, create: bool) -> Tuple[int, bool]:
        s = self.slots.get(id)
        if s is not None:
            return s, False
        if not create:
            return -1, False
        if self.kind == "range_close":
            i = int(id) - self.range_start
            if not 0 <= i < self.range_size:
                raise IndexError(f"id {id} outside the range of shard {self.shard}")
        s = len(self.keys)
        self.slots[id] = s
        self.keys.append(id)
        return s, True

    # -- the per-record server API: record, execute later ------------------------------------------
    def onPullRecv(self, id, workerPartitionIndex, ps):
        self.ops.append((0, id, workerPartitionIndex))
        if len(self.ops) >= self.max_run:
            self.flush(ps)

    def onPushRecv(self, id, deltaUpdate, ps):
        self.ops.append((1, id, deltaUpdate))
        if len(self.ops) >= self.max_run:
            self.flush(ps)

    def close(self, ps):
        self.flush(ps)
        if self.kind in ("simple_close", "loose_close", "range_close") and self.keys:
            vals = self._read_rows(list(range(len(self.keys))))
            order = range(len(self.keys))
            if self.kind == "range_close":
                order = sorted(order, key=lambda s: self.keys[s])
            for s in order:
                ps.output((self.keys[s], self.codec.decode(vals[s])))

    # -- execution -----------------------------------------------------------------------------------
    def flush(self, ps) -> None:
        ops, self.ops = self.ops, []
        i, n = 0, len(ops)
        while i < n:
            kind = ops[i][0]
            j = i
            seen = set()
            while j < n and ops[j][0] == kind:
                if kind == 1:   # a push run never holds one id twice unless the op commutes on the device
                    if ops[j][1] in seen and not self._commutes():
                        break
                    seen.add(ops[j][1])
                j += 1
            if kind == 0:
                self._run_pulls(ops[i:j], ps)
            else:
                self._run_pushes(ops[i:j], ps)
            i = j

    def _commutes(self) -> bool:
        return self.op == OP_ADD and self.kind not in ("lockA", "lockB")

    def _write_rows(self, slots: List[int], values: List[np.ndarray]) -> None:
        import torch
        from ..ops import native

        if not slots:
            return
        self._alloc(max(slots) + 1)
        with self._ctx():
            ids = torch.tensor(slots, dtype=torch.int64).to(self.device, non_blocking=True)
            vals = torch.from_numpy(np.stack(values).astype(np.float32)).to(self.device, non_blocking=True)
            self._k_assign(ids, vals.contiguous())
        self.stats["kernels"] += 1

    def _read_rows(self, slots: List[int]) -> np.ndarray:
        import torch
        from ..ops import native

        with self._ctx():
            ids = torch.tensor(slots, dtype=torch.int64).to(self.device, non_blocking=True)
            out = torch.empty((len(slots), self.codec.dim), dtype=torch.float32, device=self.device)
            self._k_gather(ids, out)
            host = out.cpu()               # stream-ordered D2H, synchronises this stream only
        self.stats["kernels"] += 1
        return host.numpy()

    def _materialise(self, ids: List[Any]) -> List[int]:
        """Slots of ``ids``, running ``paramInit`` (host closure) for the ones never seen before."""
        slots, new_slots, new_vals = [], [], []
        for id in ids:
            s, created = self._slot_of(id, True)
            if created:
                v = self.init(id)
                self.codec.learn(v)
                new_slots.append(s)
                new_vals.append(self.codec.encode(v))
            slots.append(s)
        self._write_rows(new_slots, new_vals)
        return slots

    def _run_pulls(self, run, ps) -> None:
        lock = self.kind in ("lockA", "lockB")
        if lock:
            todo = []
            for (_, id, widx) in run:
                st = self.locks.get(id)
                if st is None:
                    st = [False, deque()]
                    self.locks[id] = st
                if not st[0]:
                    st[0] = True
                    todo.append((id, widx))
                elif not (self.kind == "lockB" and widx in st[1]):
                    st[1].append(widx)
            if not todo:
                self._materialise([id for (_, id, _) in run])
                return
            run = [(0, id, w) for (id, w) in todo]
        slots = self._materialise([id for (_, id, _) in run])
        vals = self._read_rows(slots)
        self.stats["pull_runs"] += 1
        self.stats["pulls"] += len(run)
        for (_, id, widx), row in zip(run, vals):
            ps.answerPull(id, self.codec.decode(row), widx)

    def _op_for(self, value_sample, delta) -> str:
        if self.op is None or self.op == "probe-per-delta":
            scalar = isinstance(delta, (bool, int, float, np.integer, np.floating))
            if scalar and not isinstance(self.update, str):
                key = (type(delta).__name__, delta)
                op = self._op_cache.get(key)
                if op is None:
                    op = classify_update(self.update, self.codec, value_sample, delta)
                    if len(self._op_cache) < 256:
                        self._op_cache[key] = op
                self.op = "probe-per-delta"
                return op
            self.op = classify_update(self.update, self.codec, value_sample, delta)
        return self.op

    def _run_pushes(self, run, ps) -> None:
        import torch
        from ..ops import native

        lock = self.kind in ("lockA", "lockB")
        emit = self.kind in ("simple", "loose", "lockA", "lockB")
        fresh_slots, fresh_vals, fresh_out = [], [], []
        dev_ops: Dict[str, List[Tuple[int, Any, np.ndarray]]] = {}
        host_ops: List[Tuple[int, Any, Any]] = []
        for (_, id, delta) in run:
            s, created = self._slot_of(id, not lock)
            if lock and s < 0:
                raise RuntimeError("Not existed model was not able to update by any delta.")
            if created:       # push before any pull: the store-specific rule for unseen ids
                if self.kind in ("simple", "simple_close", "range_close"):
                    v = delta
                elif self.kind == "loose":
                    v = self.init(id) if self.drop_first_delta else self.update(self.init(id), delta)
                else:
                    v = self.store(delta)
                self.codec.learn(v)
                fresh_slots.append(s); fresh_vals.append(self.codec.encode(v)); fresh_out.append((id, v))
                continue
            sample = self.codec.decode(np.zeros(max(self.codec.dim, 1), dtype=np.float32))
            op = self._op_for(sample, delta)
            if op == OP_HOST:
                host_ops.append((s, id, delta))
            else:
                dev_ops.setdefault(op, []).append((s, id, ValueCodec._flat(delta)))
        self._write_rows(fresh_slots, fresh_vals)
        outputs: List[Tuple[Any, Any]] = list(fresh_out) if emit else []
        for op, items in dev_ops.items():
            slots = [s for (s, _, _) in items]
            with self._ctx():
                ids = torch.tensor(slots, dtype=torch.int64).to(self.device, non_blocking=True)
                d = torch.from_numpy(np.stack([x for (_, _, x) in items]).astype(np.float32)).to(
                    self.device, non_blocking=True).contiguous()
                if op == OP_ADD:
                    new = self._k_add(ids, d, fetch=emit)
                    new = new.cpu().numpy() if new is not None else None
                else:
                    if op == OP_ASSIGN:
                        self._k_assign(ids, d)
                        new_t = d
                    else:
                        cur = torch.empty((len(slots), self.codec.dim), dtype=torch.float32, device=self.device)
                        self._k_gather(ids, cur)
                        new_t = torch.maximum(cur, d[:, : self.codec.dim]) if op == OP_MAX else \
                            torch.minimum(cur, d[:, : self.codec.dim])
                        self._k_assign(ids, new_t.contiguous())
                    new = new_t.cpu().numpy() if emit else None
            self.stats["kernels"] += 1
            if emit:
                outputs.extend((id, self.codec.decode(row)) for (_, id, _), row in zip(items, new))
        if host_ops:          # arbitrary pure Python paramUpdate: read-modify-write of the device rows
            cur = self._read_rows([s for (s, _, _) in host_ops])
            vals = []
            for (s, id, delta), row in zip(host_ops, cur):
                v = self.update(self.codec.decode(row), delta)
                vals.append(self.codec.encode(v))
                if emit:
                    outputs.append((id, v))
            self._write_rows([s for (s, _, _) in host_ops], vals)
            self.stats["host_updates"] += len(host_ops)
        self.stats["push_runs"] += 1
        self.stats["pushes"] += len(run)
        if lock:              # hand the fresh value to the queue head (stays locked) or unlock
            handoff = []
            for (_, id, _) in run:
                st = self.locks.get(id)
                if st is None:
                    st = [False, deque()]
                    self.locks[id] = st
                if st[1]:
                    handoff.append((id, st[1].popleft()))
                    st[0] = True
                else:
                    st[0] = False
            if handoff:
                vals = self._read_rows([self.slots[id] for (id, _) in handoff])
                for (id, widx), row in zip(handoff, vals):
                    ps.answerPull(id, self.codec.decode(row), widx)
        for o in outputs:
            ps.output(o)

    # introspection (tests)
    def state(self, id):
        locked, q = self.locks.get(id, (False, deque()))
        val = self.codec.decode(self._read_rows([self.slots[id]])[0])
        return locked, val, list(q)


def torch_pad(d, width: int):
    import torch

    return d if d.shape[1] == width else torch.nn.functional.pad(d, (0, width - d.shape[1]))


def to_device_logic(psLogic):
    """Device twin of a built-in host store, or ``None`` if ``psLogic`` is user-defined server code
    (which keeps running on the host tier)."""
    from . import logics as L

    if getattr(psLogic, "is_device_store", False):
        return psLogic
    t = type(psLogic)
    if t is L.SimplePSLogic:
        return DeviceStoreLogic("simple", psLogic.init, psLogic.update)
    if t is L.SimplePSLogicWithClose:
        return DeviceStoreLogic("simple_close", psLogic.init, psLogic.update)
    if t is L.LooseSimplePSLogic:
        return DeviceStoreLogic("loose", psLogic.init, psLogic.update,
                                drop_first_delta=psLogic.drop_first_delta)
    if t is L.LooseSimplePSLogicWithClose:
        return DeviceStoreLogic("loose_close", psLogic.init, psLogic.update, store=psLogic.store)
    if t is L.RangePSLogicWithClose:
        return DeviceStoreLogic("range_close", psLogic.init, psLogic.update,
                                featureCount=psLogic.featureCount)
    if t is L.LockPSLogicA:
        return DeviceStoreLogic("lockA", psLogic.init, psLogic.update)
    if t is L.LockPSLogicB:
        return DeviceStoreLogic("lockB", psLogic.init, psLogic.update)
    return None
