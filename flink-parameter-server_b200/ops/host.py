"""ctypes bindings for ``libfps_host.so`` (native host runtime: partitioner/packer, key interner)."""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import List

import numpy as np
import torch

from . import build as _build

_lib = None
_lock = threading.Lock()


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not _build.HOST_LIB.exists():
                    _build.build_host()
                l = C.CDLL(str(_build.HOST_LIB))
                l.fps_interner_new.restype = C.c_void_p
                l.fps_interner_size.restype = C.c_int64
                l.fps_interner_size.argtypes = [C.c_void_p]
                l.fps_interner_free.argtypes = [C.c_void_p]
                _lib = l
    return _lib


def partition_pack(users: torch.Tensor, items: torch.Tensor, ratings: torch.Tensor, workers: int,
                   pin: bool = True, threads: int = 0) -> List[torch.Tensor]:
    """Bucket a block of ratings by ``user % workers`` into per-worker packed64 tensors (pinned)."""
    u = users.to(torch.int32).contiguous(); i = items.to(torch.int32).contiguous()
    r = ratings.to(torch.float32).contiguous()
    n = u.numel()
    counts = (C.c_int64 * workers)()
    lib().fps_host_count_by_worker(C.c_void_p(u.data_ptr()), C.c_int64(n), workers, counts)
    outs = [torch.empty(int(counts[w]), dtype=torch.int64, pin_memory=pin and torch.cuda.is_available())
            for w in range(workers)]
    ptrs = (C.c_void_p * workers)(*[o.data_ptr() for o in outs])
    rc = lib().fps_host_partition_pack(C.c_void_p(u.data_ptr()), C.c_void_p(i.data_ptr()),
                                       C.c_void_p(r.data_ptr()), C.c_int64(n), workers, ptrs,
                                       threads or min(8, os.cpu_count() or 1))
    if rc != 0:
        raise ValueError("ids exceed the packed64 record range (user < 2^26, item < 2^22)")
    return outs


class NativeInterner:
    """64-bit key -> dense slot, C++ hash map (thread safe)."""

    def __init__(self):
        self._p = C.c_void_p(lib().fps_interner_new())

    def map(self, keys, insert: bool = True) -> np.ndarray:
        k = np.ascontiguousarray(np.asarray(keys, dtype=np.int64))
        out = np.empty(k.shape[0], dtype=np.int32)
        lib().fps_interner_map(self._p, k.ctypes.data_as(C.c_void_p), C.c_int64(k.shape[0]),
                               out.ctypes.data_as(C.c_void_p), int(insert))
        return out

    def __len__(self) -> int:
        return int(lib().fps_interner_size(self._p))

    def keys(self) -> np.ndarray:
        out = np.empty(len(self), dtype=np.int64)
        lib().fps_interner_keys(self._p, out.ctypes.data_as(C.c_void_p))
        return out

    def __del__(self):
        try:
            lib().fps_interner_free(self._p)
        except Exception:
            pass


def mf_train(users: torch.Tensor, items: torch.Tensor, ratings: torch.Tensor, num_users: int, num_items: int,
             num_factors: int = 10, range_min: float = -0.01, range_max: float = 0.01,
             learning_rate: float = 0.01, workers: int = 4, servers: int = 4, pull_limit: int = 1600,
             epochs: int = 1, seed: int = 0, plain_residual: bool = False, negative_sample_rate: int = 0,
             user_memory: int = 128):
    """Asynchronous SGD matrix factorisation on the native host engine (``fps_host_mf_train``): worker and
    server *threads* exchanging pull / answer / push messages over lock-free SPSC rings with a pull limiter
    -- the reference's protocol at native speed, no GPU.  Returns ``(user_table, item_table, user_touched,
    item_touched, sum_sq_err)`` as numpy arrays."""
    u = users.to(torch.int32).contiguous(); i = items.to(torch.int32).contiguous()
    r = ratings.to(torch.float32).contiguous()
    n = int(u.numel())
    k = int(num_factors)
    ut = np.empty((int(num_users), k), dtype=np.float32)
    it = np.empty((int(num_items), k), dtype=np.float32)
    utouch = np.zeros(int(num_users), dtype=np.uint8)
    itouch = np.zeros(int(num_items), dtype=np.uint8)
    sse = C.c_double(0.0)
    rc = lib().fps_host_mf_train(
        C.c_void_p(u.data_ptr()), C.c_void_p(i.data_ptr()), C.c_void_p(r.data_ptr()), C.c_int64(n),
        C.c_int32(workers), C.c_int32(servers), C.c_int32(k), C.c_float(learning_rate),
        C.c_int32(1 if plain_residual else 0), C.c_float(range_min), C.c_float(range_max),
        C.c_uint64(seed & (2**64 - 1)), C.c_int32(epochs), C.c_int32(pull_limit),
        ut.ctypes.data_as(C.c_void_p), C.c_int64(num_users), it.ctypes.data_as(C.c_void_p),
        C.c_int64(num_items), utouch.ctypes.data_as(C.c_void_p), itouch.ctypes.data_as(C.c_void_p),
        C.byref(sse), C.c_int32(negative_sample_rate), C.c_int32(user_memory))
    if rc == -2:
        from ..errors import FactorIsNotANumberException

        raise FactorIsNotANumberException("non-finite SGD update in the native host engine")
    if rc != 0:
        raise ValueError("bad arguments for the native MF engine (ids out of range, k > 128, ...)")
    return ut, it, utouch.astype(bool), itouch.astype(bool), float(sse.value)


PA_ALGOS = {"PA": 0, "PAI": 1, "PAII": 2}


def _check_csr(rp: np.ndarray, c: np.ndarray, v: np.ndarray, n: int) -> None:
    """The pointers go straight to C: a malformed CSR would read out of bounds there."""
    if rp.ndim != 1 or rp.shape[0] != n + 1:
        raise ValueError(f"row_ptr must have n + 1 = {n + 1} entries, got {rp.shape}")
    if n and (rp[0] != 0 or np.any(np.diff(rp) < 0)):
        raise ValueError("row_ptr must start at 0 and be non-decreasing")
    nnz = int(rp[-1]) if rp.shape[0] else 0
    if c.shape[0] != nnz or v.shape[0] != nnz:
        raise ValueError(f"cols / vals must have row_ptr[-1] = {nnz} entries, got {c.shape[0]} / {v.shape[0]}")


def pa_binary(row_ptr, cols, vals, labels, feature_count: int, algo: str = "PA", aggressiveness: float = 0.0,
              workers: int = 4, servers: int = 4, pull_limit: int = 10000, range_partitioning: bool = False,
              weights=None):
    """Binary passive-aggressive training / prediction on the native host engine
    (``fps_host_pa_binary``).  CSR examples; ``labels`` +1 / -1 / 0 (= predict only).  ``weights``: optional
    initial model (float32 [feature_count], updated in place).  Returns ``(pred, weights, touched)``."""
    rp = np.ascontiguousarray(np.asarray(row_ptr, dtype=np.int64))
    c = np.ascontiguousarray(np.asarray(cols, dtype=np.int32))
    v = np.ascontiguousarray(np.asarray(vals, dtype=np.float32))
    y = np.ascontiguousarray(np.asarray(labels, dtype=np.int32))
    n = int(y.shape[0])
    _check_csr(rp, c, v, n)
    if weights is None:
        weights = np.zeros(int(feature_count), dtype=np.float32)
    if weights.dtype != np.float32 or not weights.flags["C_CONTIGUOUS"] or weights.shape[0] != feature_count:
        raise ValueError("weights must be a contiguous float32 array of length feature_count")
    pred = np.zeros(n, dtype=np.int32)
    touched = np.zeros(int(feature_count), dtype=np.uint8)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    rc = lib().fps_host_pa_binary(ptr(rp), ptr(c), ptr(v), ptr(y), C.c_int64(n), C.c_int64(feature_count),
                                  C.c_int32(PA_ALGOS[algo]), C.c_float(aggressiveness), C.c_int32(workers),
                                  C.c_int32(servers), C.c_int32(pull_limit), C.c_int32(1 if range_partitioning else 0),
                                  ptr(weights), ptr(pred), ptr(touched))
    if rc == -2:
        from ..errors import FactorIsNotANumberException

        raise FactorIsNotANumberException("non-finite passive-aggressive update in the native host engine")
    if rc != 0:
        raise ValueError("bad arguments for the native PA engine (feature id out of range, ...)")
    return pred, weights, touched.astype(bool)


PA_MULTI_ALGOS = {"PA": 0, "PAI": 1, "PAII": 2, "PB": 3, "ML": 4}


def pa_multiclass(row_ptr, cols, vals, labels, feature_count: int, num_labels: int, algo: str = "PA",
                  aggressiveness: float = 0.0, cost=None, workers: int = 4, servers: int = 4,
                  pull_limit: int = 10000, range_partitioning: bool = False, weights=None):
    """Multiclass passive-aggressive (one-versus-all PA / PAI / PAII, cost-based PB / ML) on the native host
    engine (``fps_host_pa_multiclass``).  ``labels`` in ``[0, L)`` or -1 (= predict only); ``cost``: optional
    ``[L, L]`` matrix (row = true label).  Returns ``(pred, weights [feature_count, L], touched)``."""
    rp = np.ascontiguousarray(np.asarray(row_ptr, dtype=np.int64))
    c = np.ascontiguousarray(np.asarray(cols, dtype=np.int32))
    v = np.ascontiguousarray(np.asarray(vals, dtype=np.float32))
    y = np.ascontiguousarray(np.asarray(labels, dtype=np.int32))
    n, L = int(y.shape[0]), int(num_labels)
    _check_csr(rp, c, v, n)
    if weights is None:
        weights = np.zeros((int(feature_count), L), dtype=np.float32)
    if weights.dtype != np.float32 or not weights.flags["C_CONTIGUOUS"] or weights.shape != (feature_count, L):
        raise ValueError("weights must be a contiguous float32 array [feature_count, num_labels]")
    cm = None if cost is None else np.ascontiguousarray(np.asarray(cost, dtype=np.float32).reshape(L, L))
    pred = np.zeros(n, dtype=np.int32)
    touched = np.zeros(int(feature_count), dtype=np.uint8)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    rc = lib().fps_host_pa_multiclass(ptr(rp), ptr(c), ptr(v), ptr(y), C.c_int64(n), C.c_int64(feature_count),
                                      C.c_int32(L), C.c_int32(PA_MULTI_ALGOS[algo]), C.c_float(aggressiveness),
                                      ptr(cm) if cm is not None else C.c_void_p(None), C.c_int32(workers),
                                      C.c_int32(servers), C.c_int32(pull_limit),
                                      C.c_int32(1 if range_partitioning else 0), ptr(weights), ptr(pred),
                                      ptr(touched))
    if rc == -2:
        from ..errors import FactorIsNotANumberException

        raise FactorIsNotANumberException("non-finite passive-aggressive update in the native host engine")
    if rc != 0:
        raise ValueError("bad arguments for the native PA engine (feature / label out of range, L > 128, ...)")
    return pred, weights, touched.astype(bool)
