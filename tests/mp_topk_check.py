"""Multi-GPU top-K serving check (psTopKGenerator capability), run under torchrun:
user vectors on the sharded PS table, every rank owns a slice of the items, each query is answered by
every rank with its local top-workerK and the partial lists are merged (E8 + E9)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from tests.mp_util import init_dist
    rank, world, dev, shared = init_dist()
    from fps_b200.models.mf.device_topk import DistributedTopK
    from fps_b200.store.sharded_table import ShardedTable

    nu, k, n_local, K = 5000, 32, 6000, 20
    users = ShardedTable(nu, k, seed=3, init_range=(-1, 1))
    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    local_items = (torch.randn(n_local, k, generator=g) * (0.2 + torch.rand(n_local, 1, generator=g))).to(dev)
    local_ids = (torch.arange(n_local, device=dev) * world + rank)
    q = torch.randint(0, nu, (300,), generator=torch.Generator().manual_seed(7)).to(dev)   # same on all ranks
    sc, ids = DistributedTopK(users, local_items, local_ids).topk(q, K)
    # reference: gather every rank's items, exact fp32 scores
    from tests.mp_util import all_gather_cat
    items = all_gather_cat(local_items); gids = all_gather_cat(local_ids)
    exact = users.pull(q) @ items.T
    ref = torch.topk(exact, K, dim=1)
    ref_ids = gids[ref.indices]
    assert sc.shape == (300, K) and ids.shape == (300, K)
    assert (sc[:, :-1] >= sc[:, 1:]).all(), "merged list not sorted"
    # TF32 scores: values close to the exact ones, sets agree except near-ties at the K-th place
    pos = (ids % world) * n_local + ids // world          # column of a global item id in the gathered table
    assert torch.equal(gids[pos], ids)
    torch.testing.assert_close(sc, torch.gather(exact, 1, pos), rtol=2e-2, atol=2e-2)
    overlap = (ids[:, :, None] == ref_ids[:, None, :]).any(-1).float().mean().item()
    assert overlap > 0.97, overlap
    assert torch.equal(ids[:, 0], ref_ids[:, 0]) or (sc[:, 0] - ref.values[:, 0]).abs().max() < 2e-2
    users.barrier()
    users.close()
    # ---- one-sided gather of partial lists (E9) and a 2-shard sketch query over it -----------------------
    from fps_b200.parallel.fabric import P2PGather
    pg = P2PGather(1024)
    for it in range(5):                      # slot reuse + acknowledgements
        parts = pg.gather(torch.full((7,), float(rank * 10 + it), device=dev), dst=None)
        assert [int(p[0]) for p in parts] == [r * 10 + it for r in range(world)]
        only0 = pg.gather(torch.arange(4, device=dev) + rank, dst=0)
        assert (only0 is None) == (rank != 0)
        if rank == 0:
            assert [int(p[0]) for p in only0] == list(range(world))
    pg.close()
    from fps_b200.models.sketch.device import DeviceSketch
    from fps_b200.models.sketch.hashing import java_string_hash
    from fps_b200.models.sketch.jobs import median_of_means
    words = [["cat", "dog"], ["cat", "dog", "fish"], ["bird"], ["bird", "fish"], ["cat", "dog"], ["fish"]] * 8
    tweets = [(2000 + i, ws) for i, ws in enumerate(words)]
    sk = DeviceSketch("tow", 64, 96)
    for w in ["cat", "dog", "fish", "bird"]:
        sk._key(w)                           # same dictionary on every rank
    sk.update(tweets[rank::world])           # every rank feeds its partition of the stream
    sk.table.barrier()
    got = sk.query("cat", 3, num_means=6)
    full = {}
    for part in [None] * 1:
        pass
    parts = [None] * world
    dist.all_gather_object(parts, sk.model())
    for p in parts:
        full.update(dict(p))
    target = full[java_string_hash("cat")]
    host = sorted(((median_of_means(v, target, 96, 6), k) for k, v in full.items()), reverse=True)[:3]
    assert {k for _, k in got[:2]} == {k for _, k in host[:2]}, (got, host)      # cat / dog tie at the top
    for (a, _), (b, _) in zip(got, host):
        assert abs(a - b) < 1e-3 * max(1.0, abs(b))
    sk.close()
    if rank == 0:
        print("MP_TOPK_CHECK_OK", round(overlap, 4))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
