"""``psOnlineMF(..., backend="native")`` / ``psOfflineMF(..., backend="native")``: the reference's
asynchronous MF protocol (PSOnlineMatrixFactorization.scala:39-75,
PSOfflineMatrixFactorizationWorker.scala:97-147) on the native host engine -- worker and server threads,
pull / answer / push messages over lock-free SPSC rings, pull limiter -- ``ops/csrc/fps_host.cpp``.

Negative sampling follows the reference worker (items seen so far by the worker, minus the user's last
``userMemory`` items, PSOnlineMatrixFactorizationWorker.scala:61-78)."""
from __future__ import annotations

import torch

from ...api import Left, Right
from ...ops import host
from ...runtime.stream import ResultStream


def ps_mf_native(src, numFactors=10, rangeMin=-0.01, rangeMax=0.01, learningRate=0.01, pullLimit=1600,
                 workerParallelism=4, psParallelism=4, seed=0, plain_residual=False, epochs=1,
                 negativeSampleRate=0, userMemory=128) -> ResultStream:
    recs = list(src.collect() if hasattr(src, "collect") else src)
    if not recs:
        return ResultStream([])
    users = torch.tensor([r.user for r in recs], dtype=torch.int32)
    items = torch.tensor([r.item for r in recs], dtype=torch.int32)
    ratings = torch.tensor([r.rating for r in recs], dtype=torch.float32)
    num_users, num_items = int(users.max()) + 1, int(items.max()) + 1
    ut, it, utouch, itouch, sse = host.mf_train(
        users, items, ratings, num_users, num_items, numFactors, rangeMin, rangeMax, learningRate,
        workers=workerParallelism, servers=psParallelism, pull_limit=max(1, int(pullLimit)), epochs=epochs,
        seed=seed, plain_residual=plain_residual, negative_sample_rate=int(negativeSampleRate),
        user_memory=int(userMemory))
    out = [Left((int(u), ut[u].astype("float64"))) for u in utouch.nonzero()[0]]
    out += [Right((int(i), it[i].astype("float64"))) for i in itouch.nonzero()[0]]
    rs = ResultStream(out)
    rs.sum_sq_err = sse
    return rs
