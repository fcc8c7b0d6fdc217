"""Multi-rank checks of the replica (sender-side combining) mode, run under torchrun:

A. conservation -- K steps x N ranks of known random pushes into the local replicas while the
   device-driven exchanges overlap them (count policy, staggered destinations, two exchanges in
   flight): after a collective refresh the master equals init + the sum of ALL deltas of ALL ranks to
   1e-5 and every replica equals the master (SimplePSLogic.scala:13-25 semantics through the batching
   senders, CombinationLogic.scala:12-33).
B. convergence gate -- the same synthetic low-rank rating stream trained with the same update budget
   by (i) one worker alone, (ii) N workers in direct one-sided mode, (iii) N workers in replica mode:
   held-out RMSE of (ii) and (iii) must match (i).
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def conservation(rank, world, dev):
    from fps_b200.ops import native
    from fps_b200.store.replica_cache import ReplicaCache
    from fps_b200.store.sharded_table import ShardedTable
    from tests.mp_util import all_reduce_sum

    for n, dim, sync_every in [(30011, 64, 2), (4099, 300, 3)]:
        t = ShardedTable(n, dim, seed=5, init_range=(-1, 1))
        all_ids = torch.arange(n, device=dev)
        init = t.pull(all_ids).clone()
        t.barrier()
        rc = ReplicaCache(t, sync_every=sync_every, max_outstanding=2)
        g = torch.Generator(device="cpu").manual_seed(17 + rank)
        total = torch.zeros(n, dim, device=dev)
        for step in range(13):
            ids = torch.randint(0, n, (6000,), generator=g).to(dev)
            delta = torch.randn(6000, dim, generator=g).to(dev)
            rc.after_step(6000)                      # policy + exchange overlap the pushes below
            native.push_add(rc.table_c, ids, delta)
            total.index_add_(0, ids, delta)
        rc.refresh()
        torch.cuda.synchronize()
        flushed = rc.flush_counts()
        assert min(flushed) >= 13 // sync_every, flushed        # every destination flushed on its count
        all_reduce_sum(total)
        torch.testing.assert_close(t.pull(all_ids), init + total, rtol=1e-5, atol=2e-5)
        torch.testing.assert_close(rc.rows(all_ids), init + total, rtol=1e-5, atol=2e-5)
        torch.testing.assert_close(rc.base, rc.cache, rtol=1e-5, atol=1e-5)   # equal up to fp32 rounding dust
        t.barrier()
        t.close()


def convergence(rank, world, dev):
    from fps_b200.models.mf.device import DeviceOnlineMF, ERR_PLAIN
    from fps_b200.utils.synthetic import lowrank_ratings
    from tests.mp_util import all_reduce_sum

    nu, ni, k, lr, init = 4096, 8192, 16, 0.05, 0.3
    batch, steps = 32768, 160 // world
    solo_groups = [dist.new_group([r]) for r in range(world)]

    def batch_of(rid, step):
        g = torch.Generator(device="cpu").manual_seed(100003 * step + rid)
        u = (torch.randint(0, nu // world, (batch,), generator=g) * world + rid).to(dev)
        i = torch.randint(0, ni, (batch,), generator=g).to(dev)
        return u.int(), i.int(), lowrank_ratings(u, i)

    gh = torch.Generator(device="cpu").manual_seed(424242)
    hu = torch.randint(0, nu, (100000,), generator=gh).to(dev)
    hi = torch.randint(0, ni, (100000,), generator=gh).to(dev)
    hr = lowrank_ratings(hu, hi)

    def rmse(model, w, r):
        mine = (hu % w) == r
        pred = model.predict(hu[mine], hi[mine])
        s = torch.stack([((hr[mine] - pred) ** 2).sum(), mine.sum().float()])
        if w > 1:
            all_reduce_sum(s)
        return float((s[0] / s[1]).sqrt())

    common = dict(range_min=-init, range_max=init, learning_rate=lr, seed=9, err_mode=ERR_PLAIN,
                  item_blocking=True, block_bytes=1 << 16)
    out = {}
    for mode, cache in (("replica", True), ("direct", False)):
        kw = dict(common)
        if not cache:
            kw["item_blocking"] = False
        m = DeviceOnlineMF(nu, ni, k, item_cache=cache, sync_every=2, **kw)
        for s in range(steps):
            m.step(*batch_of(rank, s))
        m.refresh()
        m.check_finite()
        out[mode] = rmse(m, world, rank)
        m.barrier()
        m.close()
    solo = DeviceOnlineMF(nu, ni, k, group=solo_groups[rank], **common)
    if rank == 0:
        for s in range(steps):
            for rid in range(world):
                solo.step(*batch_of(rid, s))
        ref = torch.tensor([rmse(solo, 1, 0)], device=dev)
    else:
        ref = torch.zeros(1, device=dev)
    all_reduce_sum(ref)
    out["solo"] = float(ref)
    solo.close()
    assert out["solo"] < 0.1, out                                    # the stream is learnable
    for mode in ("replica", "direct"):
        assert out[mode] <= 1.2 * out["solo"] + 0.005, out            # same quality on the same budget
    return out


def main():
    from tests.mp_util import init_dist
    rank, world, dev, shared = init_dist()
    conservation(rank, world, dev)
    q = convergence(rank, world, dev)
    dist.barrier()
    if rank == 0:
        print(f"MP_REPLICA_CHECK_OK world={world} shared_gpu={int(shared)} rmse={q}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
