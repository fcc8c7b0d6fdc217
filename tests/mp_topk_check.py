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
    if rank == 0:
        print("MP_TOPK_CHECK_OK", round(overlap, 4))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
