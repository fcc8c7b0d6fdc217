"""BASELINE.json config 1 on the CPU: asynchronous SGD matrix factorisation, k=16, world_size=2 -- the
WorkerLogic / pull / push plumbing without a GPU.  Three tiers, same synthetic rank-8 stream:

  * ``dist``   -- ``transform_distributed``: per-record Python ``WorkerLogic`` callbacks (the reference's
                  PSOnlineMatrixFactorizationWorker protocol, PSOnlineMatrixFactorizationWorker.scala:22-90),
                  one process per rank, messages over gloo.
  * ``coll``   -- the collective formulation (all_to_all of ids / rows / deltas + torch kernels) on gloo.
  * ``native`` -- the native host engine (C++ worker / server threads over SPSC rings, one process,
                  workerParallelism = psParallelism = 2).

Prints one JSON object: updates/s per tier (max over ranks of the wall time) and the training RMSE.

    python benchmarks/cpu_world2_bench.py [--ratings 200000] [--tiers dist,coll,native]
"""
from __future__ import annotations

import argparse
import json
import os
import random
import sys
import time

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

K = 16


def _stream(n, nu, ni, seed=11):
    from fps_b200.utils.synthetic import lowrank_ratings
    g = torch.Generator().manual_seed(seed)
    u = torch.randint(0, nu, (n,), generator=g)
    i = torch.randint(0, ni, (n,), generator=g)
    return u, i, lowrank_ratings(u, i)


def _rank_main(rank, world, port, a, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fps_b200.limiter import addPullLimiter
        from fps_b200.models.mf.common import Rating, vectorSum
        from fps_b200.models.mf.online import PSOnlineMatrixFactorizationWorker
        from fps_b200.parallel.nccl_baseline import NcclOnlineMF
        from fps_b200.runtime.dist_engine import transform_distributed
        from fps_b200.server.logics import SimplePSLogicWithClose

        u, i, r = _stream(a.ratings, a.users, a.items)
        sel = (u % world) == rank
        mu, mi, mr = u[sel], i[sel], r[sel]
        res = {}
        if "dist" in a.tiers:
            n = min(int(sel.sum()), a.dist_ratings // world)
            mine = [Rating(int(x), int(y), float(z)) for x, y, z in zip(mu[:n].tolist(), mi[:n].tolist(), mr[:n].tolist())]
            init = lambda id: np.random.default_rng(1000 + id).uniform(0.0, 0.3, K)
            logic = addPullLimiter(PSOnlineMatrixFactorizationWorker(K, 0.0, 0.3, 0.05, 128, 0, seed=rank + 1,
                                                                     plain_residual=True), a.pull_limit)
            dist.barrier()
            t0 = time.perf_counter()
            transform_distributed(mine, logic, SimplePSLogicWithClose(init, vectorSum),
                                  records_per_round=a.records_per_round, gather_results=False)
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64)
            cnt = torch.tensor([float(n)], dtype=torch.float64)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            dist.all_reduce(cnt)
            res["dist"] = {"updates": int(cnt.item()), "seconds": dt.item(), "updates_per_s": cnt.item() / dt.item()}
        if "coll" in a.tiers:
            m = NcclOnlineMF(a.users, a.items, K, range_min=0.0, range_max=0.3, learning_rate=0.05, seed=3,
                             err_mode=1, device=torch.device("cpu"))
            nb = mu.numel() // a.batch
            nb_t = torch.tensor([nb]); dist.all_reduce(nb_t, op=dist.ReduceOp.MIN); nb = int(nb_t.item())
            batches = [(mu[b * a.batch:(b + 1) * a.batch], mi[b * a.batch:(b + 1) * a.batch],
                        mr[b * a.batch:(b + 1) * a.batch]) for b in range(nb)]
            for b in batches[:2]:
                m.step(*b)
            dist.barrier()
            t0 = time.perf_counter()
            for _ in range(a.epochs):
                for b in batches:
                    m.step(*b)
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            err = (m.predict(mu, mi) - mr) ** 2
            tot = torch.tensor([err.sum().item(), float(err.numel())], dtype=torch.float64)
            dist.all_reduce(tot)
            n_upd = a.epochs * nb * a.batch * world
            res["coll"] = {"updates": n_upd, "seconds": dt.item(), "updates_per_s": n_upd / dt.item(),
                           "batch_per_rank": a.batch, "train_rmse": float((tot[0] / tot[1]).sqrt())}
        if rank == 0:
            q.put(res)
    finally:
        dist.destroy_process_group()


def _native(a):
    from fps_b200.ops import host
    u, i, r = _stream(a.ratings, a.users, a.items)
    args = (u.to(torch.int32), i.to(torch.int32), r.float(), a.users, a.items, K, 0.0, 0.3, 0.05)
    kw = dict(workers=2, servers=2, pull_limit=a.pull_limit, seed=3, plain_residual=True)
    host.mf_train(*args, epochs=1, **kw)          # warm-up (thread pool, page faults)
    t0 = time.perf_counter()
    ut, it, _, _, sse = host.mf_train(*args, epochs=a.epochs, **kw)
    dt = time.perf_counter() - t0
    pred = (torch.from_numpy(ut)[u.long()] * torch.from_numpy(it)[i.long()]).sum(1)
    return {"updates": a.epochs * a.ratings, "seconds": dt, "updates_per_s": a.epochs * a.ratings / dt,
            "train_rmse": float(((pred - r) ** 2).mean().sqrt()), "workers": 2, "servers": 2}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ratings", type=int, default=400_000)
    ap.add_argument("--dist-ratings", type=int, default=40_000, help="stream length of the per-record Python tier")
    ap.add_argument("--users", type=int, default=20_000)
    ap.add_argument("--items", type=int, default=5_000)
    ap.add_argument("--epochs", type=int, default=5)
    ap.add_argument("--batch", type=int, default=16_384)
    ap.add_argument("--pull-limit", type=int, default=1600)
    ap.add_argument("--records-per-round", type=int, default=1024)
    ap.add_argument("--tiers", default="dist,coll,native")
    a = ap.parse_args()
    a.tiers = a.tiers.split(",")
    out = {"config": "async SGD MF k=16, CPU, world_size=2 (BASELINE.json config 1)", "cpus": os.cpu_count(),
           "ratings": a.ratings, "users": a.users, "items": a.items, "pull_limit": a.pull_limit}
    if "dist" in a.tiers or "coll" in a.tiers:
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = 29900 + random.randrange(90)
        procs = [ctx.Process(target=_rank_main, args=(r, 2, port, a, q)) for r in range(2)]
        for p in procs:
            p.start()
        res = q.get(timeout=1800)
        for p in procs:
            p.join(timeout=60)
        out.update(res)
    if "native" in a.tiers:
        out["native"] = _native(a)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
