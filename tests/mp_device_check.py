"""Multi-GPU checks, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/mp_device_check.py

Checks the symmetric-heap fabric (every rank sees every shard), one-sided pull / push across
shards, and the fused MF step against a single-process fp32 PyTorch reference.
"""
import os

os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")  # persistent server kernel below: see parallel/rings.py
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from tests.mp_util import init_dist, all_reduce_sum, all_gather_cat
    rank, world, dev, shared = init_dist()
    from fps_b200.store.sharded_table import ShardedTable
    from fps_b200.models.mf.device import DeviceOnlineMF
    from tests.philox_ref import init_rows_ref

    # ---- fabric + pull: every rank pulls every id and must see init(id) --------------------
    n, dim = 10007, 64
    t = ShardedTable(n, dim, seed=11, init_range=(-1, 1), track_touched=True)
    ids = torch.arange(n, device=dev, dtype=torch.int64)
    got = t.pull(ids).cpu().numpy()
    ref = init_rows_ref(ids.cpu().numpy(), dim, 11, -1, 1)[:, :dim]
    assert abs(got - ref).max() < 1e-7, "cross-shard pull mismatch"
    t.barrier()
    # ---- push from every rank into every shard: table += world * delta ----------------------
    g = torch.Generator(device="cpu").manual_seed(5)
    delta = torch.randn(n, dim, generator=g).to(dev)
    t.push(ids, delta)
    t.barrier()
    got2 = t.pull(ids)
    exp = torch.from_numpy(ref).to(dev) + world * delta
    torch.testing.assert_close(got2, exp, rtol=1e-5, atol=1e-5)
    t.barrier()
    dumped_ids, _ = t.dump_local()
    assert dumped_ids.numel() == (n + world - 1 - rank) // world
    t.close()

    # ---- fused MF step, items disjoint across ranks => deterministic ------------------------
    nu, ni, k, b = 4000 * world, 3000 * world, 64, 1000
    m = DeviceOnlineMF(nu, ni, k, range_min=-0.5, range_max=0.5, learning_rate=0.05, seed=3)
    gi = torch.Generator().manual_seed(77)
    perm_items = torch.randperm(ni, generator=gi)
    my_items = perm_items[rank * b:(rank + 1) * b].to(dev)
    gu = torch.Generator().manual_seed(100 + rank)
    my_users = (torch.randperm(nu // world, generator=gu)[:b] * world + rank).to(dev)
    ratings = torch.rand(b, generator=gu).to(dev)
    V0 = m.items.pull(torch.arange(ni, device=dev))      # whole item table before
    if m.item_cache:
        assert torch.equal(m.replica.rows(torch.arange(ni, device=dev)), V0), "replica != master after init"
        assert torch.equal(m.replica.base, m.replica.cache), "base != replica after init"
    # word2vec with replica caches on both tables: one step, master must receive every delta
    from fps_b200.models.w2v import DeviceSkipGram
    sg = DeviceSkipGram(512 * world, 300, learning_rate=0.05, negative=0, seed=2)
    sg.w_out.local.uniform_(-0.05, 0.05); sg.barrier()
    if sg.rep_out is not None:                           # the replica was pulled before the uniform_ above
        from fps_b200.store.replica_cache import ReplicaCache
        sg.rep_out = ReplicaCache(sg.w_out, 4)
    Win0 = sg.w_in.pull(torch.arange(512 * world, device=dev)); Wout0 = sg.w_out.pull(torch.arange(512 * world, device=dev))
    gp = torch.Generator().manual_seed(9)
    perm = torch.randperm(512 * world, generator=gp)
    cen = perm[rank * 100:(rank + 1) * 100].to(dev); ctx = torch.randperm(512 * world, generator=gp)[rank * 100:(rank + 1) * 100].to(dev)
    sg.step(cen.int(), ctx.int()); sg.barrier()
    uu, vv = Win0[cen], Wout0[ctx]
    gg = (0.05 * (1 - torch.sigmoid((uu * vv).sum(1))))[:, None]
    dIn = torch.zeros_like(Win0).index_add_(0, cen, gg * vv); dOut = torch.zeros_like(Wout0).index_add_(0, ctx, gg * uu)
    all_reduce_sum(dIn); all_reduce_sum(dOut)
    torch.testing.assert_close(sg.w_in.pull(torch.arange(512 * world, device=dev)), Win0 + dIn, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(sg.w_out.pull(torch.arange(512 * world, device=dev)), Wout0 + dOut, rtol=1e-5, atol=1e-6)
    sg.close()
    U0 = m.users[:, :k].clone()
    m.barrier()
    m.step(my_users.int(), my_items.int(), ratings)
    m.barrier()
    V1 = m.items.pull(torch.arange(ni, device=dev))
    u = U0[my_users // world]; v = V0[my_items]
    resid = ratings - (u * v).sum(1)
    gcoef = (0.05 * torch.sigmoid(resid))[:, None]
    U_ref = U0.clone().index_add_(0, my_users // world, gcoef * v)
    torch.testing.assert_close(m.users[:, :k], U_ref, rtol=1e-5, atol=1e-6)
    dV = torch.zeros_like(V0).index_add_(0, my_items, gcoef * u)
    all_reduce_sum(dV)
    torch.testing.assert_close(V1, V0 + dV, rtol=1e-5, atol=1e-6)
    m.check_finite()
    # ---- hot contention: all ranks hammer the same few items; count conservation -----------
    m2 = DeviceOnlineMF(64 * world, 8, 16, range_min=0.1, range_max=0.2, learning_rate=0.0, seed=4)
    uu = (torch.randint(0, 64, (50000,), device=dev) * world + rank).int()
    ii = torch.randint(0, 8, (50000,), device=dev).int()
    m2.step(uu, ii, torch.ones(50000, device=dev))
    m2.barrier()
    assert m2.stats[1].item() == 50000
    m2.close(); m.close()
    # ---- message tier across GPUs: rings in peer memory, persistent server on every rank ----------
    if shared:     # two persistent server kernels would time-slice one GPU: covered by the real 2-GPU run
        dist.barrier()
        if rank == 0:
            print(f"MP_DEVICE_CHECK_OK world={world} fabric={m.items.heap.mode} shared_gpu=1")
        dist.destroy_process_group()
        return
    from fps_b200.parallel.rings import DeviceMessageServer, DeviceRingClient, RingFabric
    import time

    tab = ShardedTable(1000, 8, seed=21, init_range=(0.0, 1.0))
    ref_all = tab.pull(torch.arange(1000, device=dev)).cpu()
    rings = RingFabric(tab.stride, capacity=128)
    server = DeviceMessageServer(tab, rings, update="add")
    client = DeviceRingClient(tab, rings, pull_limit=16)
    server.start()
    dist.barrier()
    want = torch.arange(rank, 200, 3)                  # ids owned by BOTH shards, 16-credit limiter
    client.pull(want)
    got_ids, got_vals, t0 = [], [], time.time()
    while sum(x.numel() for x in got_ids) < want.numel():
        i, v = client.collect(64)
        got_ids.append(i.cpu()); got_vals.append(v.cpu())
        assert time.time() - t0 < 20, "ring answers did not arrive"
    gi, gv = torch.cat(got_ids), torch.cat(got_vals)
    order = torch.argsort(gi)
    assert torch.equal(gi[order], torch.sort(want).values)
    torch.testing.assert_close(gv[order][:, :8], ref_all[torch.sort(want).values])
    client.push(want, torch.ones(want.numel(), 8, device=dev))
    time.sleep(0.3)
    dist.barrier()
    server.stop()
    dist.barrier()
    expect = ref_all.clone()
    for r_ in range(world):
        expect[torch.arange(r_, 200, 3)] += 1
    after = tab.pull(torch.arange(1000, device=dev)).cpu()
    torch.testing.assert_close(after, expect)
    rings.close(); tab.close()
    dist.barrier()
    # ---- throughput path across GPUs: every rank runs pull->push(+1) transactions on the same 24 hot keys
    #      under LockPSLogicA; hand-overs make several server warps answer into one response ring --------------
    tab = ShardedTable(1000, 8, seed=22, init_range=(0.0, 1.0))
    ref_hot = tab.pull(torch.arange(24, device=dev)).cpu()
    rings = RingFabric(tab.stride, capacity=64, lanes=4)
    server = DeviceMessageServer(tab, rings, update="add", lock="A", pool_size=1 << 16)
    client = DeviceRingClient(tab, rings, pull_limit=128)
    dist.barrier()
    server.start()
    per_key = 150
    hot = torch.arange(24).repeat_interleave(per_key)
    hot = hot[torch.randperm(hot.numel(), generator=torch.Generator().manual_seed(50 + rank))]
    vals = client.transact(hot, torch.ones(hot.numel(), 8))
    st = client.wait()
    assert st["pulls"] == hot.numel() and st["answers"] == hot.numel() and st["credits"] == 128
    with torch.cuda.stream(client.stream):
        dist.barrier()
    client.stream.synchronize()
    server.stop()
    dist.barrier()
    seen = all_gather_cat(torch.stack([hot.to(dev).float(), vals[:, 0]], 1))     # (key, value seen) of every rank
    after = tab.pull(torch.arange(24, device=dev)).cpu()
    for k_ in range(24):
        got = torch.sort(seen[seen[:, 0] == k_, 1].cpu() - ref_hot[k_, 0]).values
        torch.testing.assert_close(got, torch.arange(world * per_key, dtype=torch.float32), rtol=0, atol=2e-3)
    torch.testing.assert_close(after[:, 0], ref_hot[:, 0] + world * per_key, rtol=0, atol=2e-3)
    rings.close(); tab.close()
    dist.barrier()
    if rank == 0:
        print(f"MP_DEVICE_CHECK_OK world={world} fabric={m.items.heap.mode}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
