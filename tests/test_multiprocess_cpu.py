"""BASELINE.json config 1: async SGD matrix factorisation k=16 on CPU, world_size=2 (gloo) -- the
WorkerLogic / pull / push plumbing without a GPU, plus the collective (all_to_all) MF path."""
import os
import random

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _ratings(seed=47, n=160, nu=24, ni=18):
    r = random.Random(seed)
    seen, out = set(), []
    while len(out) < n:
        u, i = r.randrange(nu), r.randrange(ni)
        if (u, i) not in seen:
            seen.add((u, i)); out.append((u, i, r.random()))
    return out


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fps_b200.limiter import addPullLimiter
        from fps_b200.models.mf.common import Rating, vectorSum
        from fps_b200.models.mf.online import PSOnlineMatrixFactorizationWorker
        from fps_b200.parallel.nccl_baseline import NcclOnlineMF
        from fps_b200.runtime.dist_engine import transform_distributed
        from fps_b200.server.logics import SimplePSLogicWithClose

        data = _ratings()
        mine = [Rating(u, i, x) for (u, i, x) in data if u % world == rank] * 40   # 40 passes over the data
        rng = np.random.default_rng(5)
        init = lambda i: np.random.default_rng(1000 + i).uniform(0.0, 0.3, 16)
        logic = addPullLimiter(PSOnlineMatrixFactorizationWorker(16, 0.0, 0.3, 0.05, 128, 0, seed=rank + 1,
                                                                 plain_residual=True), 20)
        out = transform_distributed(mine, logic, SimplePSLogicWithClose(init, vectorSum))
        users, items = {}, {}
        for u, v in out.worker_outputs():
            users[u] = v
        for i, v in out.ps_outputs():
            items[i] = v
        rmse = (sum((x - float(np.dot(users[u], items[i]))) ** 2 for u, i, x in data) / len(data)) ** 0.5
        # collective path (all_to_all + torch kernels) on the same data
        m = NcclOnlineMF(24, 18, 16, range_min=0.0, range_max=0.3, learning_rate=0.05, seed=3, err_mode=1,
                         device=torch.device("cpu"))
        u = torch.tensor([d[0] for d in data if d[0] % world == rank], dtype=torch.int64)
        it = torch.tensor([d[1] for d in data if d[0] % world == rank], dtype=torch.int64)
        r = torch.tensor([d[2] for d in data if d[0] % world == rank], dtype=torch.float32)
        for _ in range(150):
            m.step(u, it, r)
        err = (m.predict(u, it) - r) ** 2
        tot = torch.tensor([err.sum().item(), float(err.numel())])
        dist.all_reduce(tot)
        if rank == 0:
            q.put((rmse, (tot[0] / tot[1]).sqrt().item(), len(items)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_cpu_plumbing_and_collective_mf():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + random.randrange(200)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    rmse_ps, rmse_coll, n_items = q.get(timeout=5)
    assert n_items == 18
    assert rmse_ps <= 0.5, rmse_ps          # reference quality gate (RMSE <= 0.5)
    assert rmse_coll <= 0.5, rmse_coll
