"""Device message tier: peer-memory rings + persistent server kernel + credit-counter client.

See ops/csrc/fps_rings.cu.  ``RingFabric`` lays the rings out in one symmetric-heap allocation
per rank; ``DeviceMessageServer`` runs the persistent ``fps_server_loop`` kernel for one shard;
``DeviceRingClient`` is the worker-side ``ParameterServerClient`` equivalent working on id tensors,
with the pull limiter implemented as a device-resident credit counter + FIFO spill queue.

While a server kernel is resident never call ``torch.cuda.synchronize()`` (device-wide sync waits
for the persistent kernel); synchronise streams / events instead.  Run such processes with
``CUDA_MODULE_LOADING=EAGER`` (set before CUDA initialises): CUDA otherwise loads kernels lazily and the
first launch of a not-yet-loaded kernel synchronises the context -- a deadlock behind a resident
kernel.  ``start()`` pre-loads this library's kernels and the torch kernels the server / client use
themselves, and warns when eager loading is not selected.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from ..ops import native
from ..store.sharded_table import ShardedTable
from .fabric import SymmetricHeap

RING_MAX_PEERS = 16
OP_PULL, OP_PUSH = 1, 2
UPDATE_OPS = {"add": 0, "assign": 1, "max": 2, "min": 3}
LOCK_MODES = {None: 0, "none": 0, "A": 1, "B": 2}
ERRORS = {1: "spin limit exceeded (peer not draining its ring)", 2: "push to an unknown id (LockPSLogic)",
          3: "waiter pool / spill queue exhausted"}


class RingSetC(C.Structure):
    _fields_ = [("base", C.c_void_p), ("n_peers", C.c_int), ("lanes", C.c_int), ("capacity", C.c_int),
                ("stride", C.c_int), ("entry_bytes", C.c_int), ("pad_", C.c_int)]


class TxnArgsC(C.Structure):
    _fields_ = [("req", RingSetC), ("resp", RingSetC), ("ids", C.c_void_p), ("deltas", C.c_void_p),
                ("seg", C.c_void_p), ("out_vals", C.c_void_p), ("credits", C.c_void_p), ("err", C.c_void_p),
                ("counters", C.c_void_p), ("self", C.c_int), ("mode", C.c_int), ("per_ring_cap", C.c_int),
                ("pad_", C.c_int)]


TXN_PULL_PUSH, TXN_PULL_ONLY, TXN_PUSH_ONLY = 0, 1, 2


class ServerArgsC(C.Structure):
    _fields_ = [("req", RingSetC), ("resp", RingSetC), ("tab", native.ShardTableC), ("self", C.c_int),
                ("update_op", C.c_int), ("lock_mode", C.c_int), ("lock_state", C.c_void_p),
                ("lock_mutex", C.c_void_p), ("wait_head", C.c_void_p), ("pool", C.c_void_p),
                ("pool_next", C.c_void_p), ("free_head", C.c_void_p), ("pool_size", C.c_int),
                ("touched", C.c_void_p), ("stop", C.c_void_p), ("err", C.c_void_p),
                ("counters", C.c_void_p), ("resp_reserve", C.c_void_p), ("resp_published", C.c_void_p),
                ("resp_tail_cache", C.c_void_p)]


class ClientArgsC(C.Structure):
    _fields_ = [("req", RingSetC), ("resp", RingSetC), ("tab", native.ShardTableC), ("st", C.c_void_p),
                ("spill", C.c_void_p), ("spill_cap", C.c_int), ("self", C.c_int)]


class RingFabric:
    """Request + response rings of every (worker, shard) pair; rank r is worker r and shard r.

    ``lanes`` parallel rings per pair (a key always travels on lane ``slot(id) % lanes``, which keeps the
    per-key FIFO order the reference's answer queues rely on): the persistent server serves every ring
    with its own warp, spread over as many CTAs as needed -- the throughput knob of the message tier."""

    def __init__(self, stride: int, capacity: int = 1024, group=None, device: Optional[int] = None,
                 lanes: int = 1):
        assert capacity & (capacity - 1) == 0, "ring capacity must be a power of two"
        ready = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if ready else 1
        self.rank = dist.get_rank(group) if ready else 0
        self.stride, self.capacity, self.lanes = int(stride), int(capacity), max(1, int(lanes))
        lib = native.lib()
        self.entry_bytes = lib.fps_ring_entry_bytes(self.stride)
        self.ring_bytes = (lib.fps_ring_bytes(self.capacity, self.stride) + 255) // 256 * 256
        # per rank: request rings [worker][lane], then response rings [shard][lane]
        self.heap = SymmetricHeap(2 * self.world * self.lanes * self.ring_bytes, group=group, device=device)
        self._tables = []        # device pointer tables (kept alive)

    def _set(self, bases) -> RingSetC:
        dev = torch.device("cuda", self.heap.device)
        tab = torch.tensor(list(bases), dtype=torch.int64, device=dev)
        self._tables.append(tab)
        r = RingSetC()
        r.base = tab.data_ptr()
        r.n_peers, r.lanes, r.capacity = self.world, self.lanes, self.capacity
        r.stride, r.entry_bytes = self.stride, self.entry_bytes
        return r

    def _off(self, block: int, peer: int, lane: int) -> int:
        return ((block * self.world + peer) * self.lanes + lane) * self.ring_bytes

    # server view: my request rings (local), the workers' response rings for me (peer)
    def server_sets(self) -> Tuple[RingSetC, RingSetC]:
        W, L, me = self.world, self.lanes, self.rank
        req = self._set([self.heap.local_ptr + self._off(0, w, l) for w in range(W) for l in range(L)])
        resp = self._set([self.heap.peer_ptrs[w] + self._off(1, me, l) for w in range(W) for l in range(L)])
        return req, resp

    # client view: every shard's request ring for me (peer), my response rings (local)
    def client_sets(self) -> Tuple[RingSetC, RingSetC]:
        W, L, me = self.world, self.lanes, self.rank
        req = self._set([self.heap.peer_ptrs[s] + self._off(0, me, l) for s in range(W) for l in range(L)])
        resp = self._set([self.heap.local_ptr + self._off(1, s, l) for s in range(W) for l in range(L)])
        return req, resp

    def close(self):
        self.heap.close()


class DeviceMessageServer:
    """Persistent server kernel for the local shard of ``table``."""

    def __init__(self, table: ShardedTable, rings: RingFabric, update: str = "add",
                 lock: Optional[str] = None, pool_size: int = 1 << 16, require_pull_before_push=None):
        self.table, self.rings = table, rings
        dev = table.cuda_device
        rows = table.rows_per_shard
        self.lock = LOCK_MODES[lock]
        z = lambda n, dt=torch.int32: torch.zeros(n, dtype=dt, device=dev)
        self.lock_state, self.lock_mutex = z(rows), z(rows)
        self.wait_head = torch.full((rows,), -1, dtype=torch.int32, device=dev)
        self.pool = z(3 * pool_size)
        self.pool_next, self.free_head = z(1), torch.full((1,), -1, dtype=torch.int32, device=dev)
        track = self.lock != 0 if require_pull_before_push is None else require_pull_before_push
        self.touched = z((rows + 31) // 32) if track else None
        self.stop_flag, self.err = z(1), z(1)
        self.counters = z(3, torch.int64)
        # multi-producer response rings: reservation / in-order publish counters (fresh rings start at 0)
        self.resp_reserve = z(rings.world * rings.lanes, torch.int64)
        self.resp_published = z(rings.world * rings.lanes, torch.int64)
        self.resp_tail_cache = z(rings.world * rings.lanes, torch.int64)
        a = ServerArgsC()
        a.req, a.resp = rings.server_sets()
        a.tab = table.table_c
        a.self, a.update_op, a.lock_mode = rings.rank, UPDATE_OPS[update], self.lock
        a.lock_state, a.lock_mutex = self.lock_state.data_ptr(), self.lock_mutex.data_ptr()
        a.wait_head, a.pool = self.wait_head.data_ptr(), self.pool.data_ptr()
        a.pool_next, a.free_head, a.pool_size = self.pool_next.data_ptr(), self.free_head.data_ptr(), pool_size
        a.touched = self.touched.data_ptr() if self.touched is not None else None
        a.stop, a.err, a.counters = self.stop_flag.data_ptr(), self.err.data_ptr(), self.counters.data_ptr()
        a.resp_reserve, a.resp_published = self.resp_reserve.data_ptr(), self.resp_published.data_ptr()
        a.resp_tail_cache = self.resp_tail_cache.data_ptr()
        self.args = a
        self.stream = torch.cuda.Stream(device=dev)
        self.ctl = torch.cuda.Stream(device=dev)
        self.running = False

    def start(self) -> None:
        import os
        import warnings

        if os.environ.get("CUDA_MODULE_LOADING", "").upper() != "EAGER":
            warnings.warn("persistent server kernel without CUDA_MODULE_LOADING=EAGER: any kernel first "
                          "launched while the server is resident can deadlock (see parallel/rings.py)")
        # CUDA loads kernels lazily and the first launch of an unloaded kernel synchronises the
        # context -- a deadlock once the persistent kernel is resident.  Load everything used while
        # the server runs *now*: our ring kernels, and the torch kernels of stop()/stats()/clients.
        native._check(native.lib().fps_rings_preload(), "rings_preload")
        with torch.cuda.stream(self.ctl):
            self.stop_flag.fill_(0)
            _ = self.counters.to("cpu"); _ = self.err.to("cpu")
            w = torch.zeros(4, dtype=torch.int64, device=self.stop_flag.device)
            _ = (w + 1).to(torch.float32).contiguous(); _ = torch.empty(4, device=w.device).zero_()
        self.ctl.synchronize()
        torch.cuda.current_stream().synchronize()
        native._check(native.lib().fps_server_loop_launch(C.byref(self.args), C.c_void_p(self.stream.cuda_stream)),
                      "server_loop_launch")
        native._bump()
        self.running = True

    def stop(self) -> None:
        if not self.running:
            return
        with torch.cuda.stream(self.ctl):
            self.stop_flag.fill_(1)
        self.ctl.synchronize()
        self.stream.synchronize()
        self.running = False
        code = int(self._read(self.err)[0])
        if code:
            raise RuntimeError(f"device server: {ERRORS.get(code, code)}")

    def stop_flag_only(self) -> None:
        """Ask the kernel to exit without waiting or raising (error paths)."""
        with torch.cuda.stream(self.ctl):
            self.stop_flag.fill_(1)
        self.ctl.synchronize()
        self.running = False

    def _read(self, t: torch.Tensor) -> torch.Tensor:
        with torch.cuda.stream(self.ctl):
            out = t.to("cpu", non_blocking=False)
        return out

    def stats(self) -> dict:
        c = self._read(self.counters).tolist()
        return {"pulls": c[0], "pushes": c[1], "answers": c[2]}


class DeviceRingClient:
    """Worker-side client: ``pull(ids)`` / ``push(ids, deltas)`` / ``collect()`` on id tensors with a
    device credit counter of ``pull_limit`` unanswered pulls."""

    def __init__(self, table: ShardedTable, rings: RingFabric, pull_limit: int = 1600,
                 spill_capacity: int = 1 << 16):
        dev = table.cuda_device
        self.dev, self.stride = dev, rings.stride
        self.table, self.rings = table, rings
        self.pull_limit = int(pull_limit)
        self.txn_credits = torch.tensor([pull_limit, 0], dtype=torch.int32, device=dev)
        self.txn_err = torch.zeros(1, dtype=torch.int32, device=dev)
        self.txn_counters = torch.zeros(3, dtype=torch.int64, device=dev)
        self.state = torch.zeros(10, dtype=torch.int32, device=dev)
        self.state[0] = pull_limit
        self.state[1] = pull_limit
        self.spill = torch.zeros(spill_capacity, dtype=torch.int64, device=dev)
        a = ClientArgsC()
        a.req, a.resp = rings.client_sets()
        a.tab = table.table_c
        a.st, a.spill, a.spill_cap, a.self = self.state.data_ptr(), self.spill.data_ptr(), spill_capacity, rings.rank
        self.args = a
        self.stream = torch.cuda.Stream(device=dev)
        self.n_out = torch.zeros(1, dtype=torch.int32, device=dev)
        # pre-load the torch kernels this client launches later (see DeviceMessageServer.start)
        with torch.cuda.stream(self.stream):
            _ = torch.arange(4).to(dev, torch.int64).contiguous()
            d = torch.zeros((2, self.stride), dtype=torch.float32, device=dev)
            d[:, :1] = torch.ones(2, 1, device=dev)
            _ = torch.empty(4, dtype=torch.int64, device=dev)[:2].to("cpu")
            _ = self.state.to("cpu"); _ = self.n_out.to("cpu")
            if rings.lanes >= 1:     # warm the torch kernels of the batched path (lazy loading, see start())
                t = torch.arange(64, device=dev, dtype=torch.int64)
                r_ = (t % 3) * 2 + (t // 3) % 2
                o_ = torch.argsort(r_, stable=True)
                sg = torch.zeros(8, dtype=torch.int32, device=dev)
                sg[1:] = torch.cumsum(torch.bincount(r_, minlength=7), 0).to(torch.int32)
                dd = torch.zeros((64, self.stride), dtype=torch.float32, device=dev)
                dd[:, :1] = torch.ones(64, 1, device=dev)[o_]
                oo = torch.empty_like(dd); oo[o_] = dd
                _ = torch.clamp(t // 5, max=3); _ = t[o_].contiguous()
                _ = self.txn_err.to("cpu"); _ = self.txn_counters.to("cpu"); _ = self.txn_credits.to("cpu")
        self.stream.synchronize()
        torch.cuda.current_stream().synchronize()

    def _issue(self, ids: torch.Tensor, deltas: Optional[torch.Tensor], op: int) -> None:
        ids = ids.to(self.dev, torch.int64).contiguous()
        with torch.cuda.stream(self.stream):
            native._check(native.lib().fps_client_issue(
                C.byref(self.args), C.c_void_p(ids.data_ptr()),
                C.c_void_p(deltas.data_ptr() if deltas is not None else None), int(ids.numel()), op,
                C.c_void_p(self.stream.cuda_stream)), "client_issue")
            native._bump()
        self.stream.synchronize()

    def pull(self, ids: torch.Tensor) -> None:
        self._issue(ids, None, OP_PULL)

    def push(self, ids: torch.Tensor, deltas: torch.Tensor) -> None:
        d = torch.zeros((ids.numel(), self.stride), dtype=torch.float32, device=self.dev)
        d[:, : deltas.shape[1]] = deltas
        self._issue(ids, d, OP_PUSH)

    def collect(self, max_n: int = 1024) -> Tuple[torch.Tensor, torch.Tensor]:
        """Consume up to ``max_n`` answers (= ``onPullRecv`` calls); releases credits and issues queued pulls."""
        ids = torch.empty(max_n, dtype=torch.int64, device=self.dev)
        vals = torch.empty((max_n, self.stride), dtype=torch.float32, device=self.dev)
        with torch.cuda.stream(self.stream):
            native._check(native.lib().fps_client_collect(
                C.byref(self.args), C.c_void_p(ids.data_ptr()), C.c_void_p(vals.data_ptr()), int(max_n),
                C.c_void_p(self.n_out.data_ptr()), C.c_void_p(self.stream.cuda_stream)), "client_collect")
            native._bump()
            n = int(self.n_out.to("cpu")[0])
        return ids[:n], vals[:n]

    # ---- throughput path: a whole micro-batch through ONE persistent client kernel ------------------------
    def _txn(self, ids: torch.Tensor, deltas: Optional[torch.Tensor], mode: int):
        dev, stride = self.dev, self.stride
        ids = ids.to(dev, torch.int64).contiguous()
        n = ids.numel()
        tab, L = self.table, self.rings.lanes
        with torch.cuda.stream(self.stream):
            if tab.mode == native.PART_HASH:
                owner, slot = ids % tab.n_shards, ids // tab.n_shards
            else:
                owner = torch.clamp(ids // tab.div, max=tab.n_shards - 1)
                slot = ids - owner * tab.div
            ring = owner * L + slot % L
            order = torch.argsort(ring, stable=True)          # per-key order is kept inside a ring
            seg = torch.zeros(tab.n_shards * L + 1, dtype=torch.int32, device=dev)
            seg[1:] = torch.cumsum(torch.bincount(ring, minlength=tab.n_shards * L), 0).to(torch.int32)
            sid = ids[order].contiguous()
            d = None
            if deltas is not None:
                d = torch.zeros((n, stride), dtype=torch.float32, device=dev)
                d[:, : deltas.shape[1]] = deltas.to(dev)[order]
            out = torch.zeros((n, stride), dtype=torch.float32, device=dev) if mode != TXN_PUSH_ONLY else None
            a = TxnArgsC()
            a.req, a.resp = self.args.req, self.args.resp
            a.ids, a.seg = sid.data_ptr(), seg.data_ptr()
            a.deltas = d.data_ptr() if d is not None else None
            a.out_vals = out.data_ptr() if out is not None else None
            a.credits, a.err, a.counters = self.txn_credits.data_ptr(), self.txn_err.data_ptr(), self.txn_counters.data_ptr()
            a.self, a.mode = self.rings.rank, mode
            n_rings = tab.n_shards * L
            a.per_ring_cap = max(32, -(-self.pull_limit // n_rings))     # fair share of the credit pool
            native._check(native.lib().fps_client_txn(C.byref(a), C.c_void_p(self.stream.cuda_stream)), "client_txn")
            native._bump()
            res = None
            if out is not None:
                res = torch.empty_like(out)
                res[order] = out
        self._keep = (sid, seg, d, out)
        return res

    def transact(self, ids: torch.Tensor, deltas: torch.Tensor) -> torch.Tensor:
        """For every key: pull, and on the answer push ``delta`` (``onPullRecv -> ps.push``), all inside one
        persistent kernel under the device credit counter.  Returns the pulled values ``[n, stride]`` (a
        future on the client stream: call :meth:`wait` before reading)."""
        return self._txn(ids, deltas, TXN_PULL_PUSH)

    def pull_all(self, ids: torch.Tensor) -> torch.Tensor:
        return self._txn(ids, None, TXN_PULL_ONLY)

    def push_all(self, ids: torch.Tensor, deltas: torch.Tensor) -> None:
        self._txn(ids, deltas, TXN_PUSH_ONLY)

    def wait(self) -> dict:
        self.stream.synchronize()
        with torch.cuda.stream(self.stream):
            code = int(self.txn_err.to("cpu")[0])
            c = self.txn_counters.to("cpu").tolist()
            cr = self.txn_credits.to("cpu").tolist()
        if code:
            raise RuntimeError(f"device client: {ERRORS.get(code, code)}")
        return {"pulls": c[0], "pushes": c[1], "answers": c[2], "credits": cr[0], "stalls": cr[1]}

    def counters(self) -> dict:
        with torch.cuda.stream(self.stream):
            s = self.state.to("cpu").tolist()
        code = s[9]
        if code:
            raise RuntimeError(f"device client: {ERRORS.get(code, code)}")
        return {"credits": s[0], "limit": s[1], "issued": s[2] | (s[3] << 32),
                "queued": (s[6] | (s[7] << 32)) - (s[4] | (s[5] << 32))}
