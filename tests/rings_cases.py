"""Device message tier: rings + persistent server kernel + credit counter.
Contracts ported from T/WorkerLogicTest.scala:34-46 and T/server/LockPSLogic{A,B}Test.scala."""
import time

import pytest
import torch

pytestmark = pytest.mark.gpu


def _collect_exact(client, n, timeout=10.0):
    ids, vals, t0 = [], [], time.time()
    while sum(x.numel() for x in ids) < n:
        i, v = client.collect(n - sum(x.numel() for x in ids))
        ids.append(i.cpu()); vals.append(v.cpu())
        assert time.time() - t0 < timeout, "answers did not arrive"
    return torch.cat(ids), torch.cat(vals)


def _setup(dim=8, n=1000, update="add", lock=None, limit=1600):
    from fps_b200.parallel.rings import DeviceMessageServer, DeviceRingClient, RingFabric
    from fps_b200.store.sharded_table import ShardedTable

    torch.cuda.set_device(0)
    table = ShardedTable(n, dim, seed=9, init_range=(0.0, 1.0))
    rings = RingFabric(table.stride, capacity=256)
    server = DeviceMessageServer(table, rings, update=update, lock=lock)
    client = DeviceRingClient(table, rings, pull_limit=limit)
    server.start()
    return table, rings, server, client


def test_device_credit_counter_matches_pull_limiter_contract():
    table, rings, server, client = _setup(limit=10)
    try:
        client.pull(torch.arange(1, 21))
        assert client.counters()["issued"] == 10 and client.counters()["queued"] == 10
        ids, _ = _collect_exact(client, 5)
        assert ids.tolist() == [1, 2, 3, 4, 5]            # FIFO per (worker, shard) pair
        assert client.counters()["issued"] == 15
        client.pull(torch.tensor([21]))
        assert client.counters()["issued"] == 15
        ids, _ = _collect_exact(client, 16)
        assert ids.tolist() == list(range(6, 22))
        c = client.counters()
        assert c["issued"] == 21 and c["credits"] == 10 and c["queued"] == 0
    finally:
        server.stop(); rings.close(); table.close()


@pytest.mark.parametrize("update", ["add", "assign", "max"])
def test_pull_push_through_rings_with_registered_update_ops(update):
    table, rings, server, client = _setup(update=update)
    try:
        ref = table.local[:, :8].clone()
        ids = torch.tensor([3, 7, 3, 500])
        client.pull(ids)
        got_ids, got = _collect_exact(client, 4)
        assert got_ids.tolist() == ids.tolist()
        torch.testing.assert_close(got[:, :8], ref[ids].cpu())
        d = torch.full((4, 8), 0.5, device="cuda"); d[2] = 2.0
        client.push(ids, d)
        client.pull(torch.tensor([3, 7]))
        _, after = _collect_exact(client, 2)
        if update == "add":
            exp3, exp7 = ref[3] + 2.5, ref[7] + 0.5
        elif update == "assign":
            exp3, exp7 = torch.full((8,), 2.0), torch.full((8,), 0.5)
        else:
            exp3, exp7 = torch.maximum(ref[3].cpu(), torch.tensor(2.0)), torch.maximum(ref[7].cpu(), torch.tensor(0.5))
        torch.testing.assert_close(after[0, :8], exp3.cpu()); torch.testing.assert_close(after[1, :8], exp7.cpu())
        assert server.stats() == {"pulls": 6, "pushes": 4, "answers": 6}
    finally:
        server.stop(); rings.close(); table.close()


@pytest.mark.parametrize("mode,waiters_kept", [("A", 2), ("B", 1)])
def test_lock_logic_on_device(mode, waiters_kept):
    table, rings, server, client = _setup(lock=mode)
    try:
        client.pull(torch.tensor([42]))
        ids, v0 = _collect_exact(client, 1)               # first pull answered, lock taken
        client.pull(torch.tensor([42, 42]))               # both queue behind the lock (B: deduplicated)
        time.sleep(0.05)
        assert client.collect(4)[0].numel() == 0
        delta = torch.ones(1, 8, device="cuda")
        client.push(torch.tensor([42]), delta)            # hand over to the queue head, stay locked
        ids, v1 = _collect_exact(client, 1)
        torch.testing.assert_close(v1[0, :8], v0[0, :8] + 1)
        extra = 0
        for _ in range(3):                                # drain the remaining waiters
            client.push(torch.tensor([42]), delta)
            time.sleep(0.02)
            extra += client.collect(4)[0].numel()
        assert extra == waiters_kept - 1
        client.pull(torch.tensor([42]))                   # unlocked again -> answered immediately
        _, v2 = _collect_exact(client, 1)
        torch.testing.assert_close(v2[0, :8], v0[0, :8] + 4)
    finally:
        server.stop(); rings.close(); table.close()


def test_lock_push_before_pull_is_an_error():
    table, rings, server, client = _setup(lock="A")
    client.push(torch.tensor([5]), torch.ones(1, 8, device="cuda"))
    time.sleep(0.05)
    with pytest.raises(RuntimeError):
        server.stop()
    rings.close(); table.close()


def test_transform_rings_runs_reference_style_worker_logic_against_device_server():
    """WorkerLogic callbacks on the host, the store (+ lock logic) in the persistent server kernel."""
    from fps_b200 import WorkerLogic
    from fps_b200.runtime.ring_engine import transform_rings
    from fps_b200.store.sharded_table import ShardedTable

    torch.cuda.set_device(0)

    class PullThenPushOne(WorkerLogic):
        def onRecv(self, data, ps):
            ps.pull(data)

        def onPullRecv(self, paramId, paramValue, ps):
            ps.output((paramId, float(paramValue[0])))
            ps.push(paramId, torch.ones(4))

    table = ShardedTable(64, 4, init="zeros")
    data = [i % 16 for i in range(48)]            # every key pulled 3 times, +1 per answer
    out = transform_rings(data, PullThenPushOne(), table, update="add", lock="A", pull_limit=8)
    model = {i: v for i, v in out.ps_outputs()}
    assert sorted(model) == list(range(16))        # lock mode tracks "exists": only pulled keys dumped
    for i in range(16):
        assert torch.equal(model[i], torch.full((4,), 3.0))
    # with per-key locks every worker sees a serialised counter: values 0, 1, 2 per key
    seen = {}
    for i, v in out.worker_outputs():
        seen.setdefault(i, []).append(v)
    assert all(sorted(v) == [0.0, 1.0, 2.0] for v in seen.values())
    assert out.server_stats["pulls"] == 48 and out.server_stats["pushes"] == 48
    assert out.client_counters["issued"] == 48 and out.client_counters["credits"] == 8
    table.close()


# ---- throughput path: multi-lane rings, multi-CTA server, one persistent client kernel per batch ----------------
def _setup_mt(dim=8, n=4000, update="add", lock=None, limit=256, lanes=8, capacity=128):
    from fps_b200.parallel.rings import DeviceMessageServer, DeviceRingClient, RingFabric
    from fps_b200.store.sharded_table import ShardedTable

    torch.cuda.set_device(0)
    table = ShardedTable(n, dim, seed=9, init_range=(0.0, 1.0))
    rings = RingFabric(table.stride, capacity=capacity, lanes=lanes)
    server = DeviceMessageServer(table, rings, update=update, lock=lock)
    client = DeviceRingClient(table, rings, pull_limit=limit)
    server.start()
    return table, rings, server, client


def test_batched_pull_then_push_transactions_on_many_lanes():
    table, rings, server, client = _setup_mt()
    try:
        ref = table.local[:, :8].clone()
        g = torch.Generator().manual_seed(1)
        ids = torch.randperm(4000, generator=g)[:3000]
        deltas = torch.randn(3000, 8, generator=g)
        vals = client.transact(ids, deltas)
        st = client.wait()
        assert st["pulls"] == 3000 and st["pushes"] == 3000 and st["answers"] == 3000
        assert st["credits"] == 256                                   # every credit returned
        torch.testing.assert_close(vals[:, :8].cpu(), ref[ids].cpu())  # unique keys: the value before the push
        server.stop()
        exp = ref.clone(); exp[ids.cuda()] += deltas.cuda()
        torch.testing.assert_close(table.local[:, :8], exp)
        assert server.stats() == {"pulls": 3000, "pushes": 3000, "answers": 3000}
    finally:
        server.stop(); rings.close(); table.close()


def test_lock_a_serialises_read_modify_write_of_a_hot_key():
    """20 keys x 200 transactions each (pull, then push +1 on the answer) under LockPSLogicA: the lock makes
    every pull see a distinct committed value, so the answers of a key are exactly init, init+1, ..."""
    table, rings, server, client = _setup_mt(lock="A", limit=512, lanes=4)
    try:
        ref = table.local[:, :8].clone()
        ids = torch.arange(20).repeat_interleave(200)
        ids = ids[torch.randperm(ids.numel(), generator=torch.Generator().manual_seed(2))]
        vals = client.transact(ids, torch.ones(ids.numel(), 8))
        client.wait()
        server.stop()
        for k in range(20):
            seen = torch.sort(vals[ids == k, 0].cpu() - ref[k, 0].cpu()).values
            torch.testing.assert_close(seen, torch.arange(200, dtype=torch.float32), rtol=0, atol=1e-3)
        torch.testing.assert_close(table.local[:20, 0], ref[:20, 0] + 200)
    finally:
        server.stop(); rings.close(); table.close()


def test_push_only_stream_applies_a_non_commutative_assign_in_key_order():
    table, rings, server, client = _setup_mt(update="assign", lanes=8)
    try:
        ids = torch.arange(50).repeat(40)                      # 40 assignments per key, in stream order
        vals = torch.arange(ids.numel(), dtype=torch.float32)[:, None].expand(-1, 8).contiguous()
        client.push_all(ids, vals)
        st = client.wait()
        assert st["pushes"] == ids.numel()
        server.stop()
        last = torch.arange(50, dtype=torch.float32) + 39 * 50    # the LAST assignment of every key wins
        torch.testing.assert_close(table.local[:50, 0].cpu(), last)
    finally:
        server.stop(); rings.close(); table.close()
