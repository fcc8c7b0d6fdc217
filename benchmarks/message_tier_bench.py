"""Message-tier throughput: pull -> push transactions through peer-memory rings served by the persistent
multi-CTA server kernel (logic-bearing stores: LockPSLogicA, non-commutative ``assign``).

    python benchmarks/message_tier_bench.py                          # one GPU (rings in local HBM)
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 benchmarks/message_tier_bench.py

Every rank is worker r and shard r; keys are uniform over the table, so (N-1)/N of the messages cross
NVLink.  One "message" = one ring entry processed by a server warp (a pull or a push; answers are not
counted).  Device-timed on the client stream, max over ranks; prints one JSON line."""
import argparse
import json
import os
import sys

os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")      # persistent kernel: see parallel/rings.py
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--keys", type=int, default=1 << 20, help="table size")
    p.add_argument("--txn", type=int, default=1 << 19, help="transactions per batch per rank")
    p.add_argument("--dim", type=int, default=8)
    p.add_argument("--lanes", type=int, default=16)
    p.add_argument("--capacity", type=int, default=256)
    p.add_argument("--limit", type=int, default=16384)
    p.add_argument("--iters", type=int, default=3)
    a = p.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    lr_ = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr_); dev = torch.device("cuda", lr_)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from fps_b200.parallel.rings import DeviceMessageServer, DeviceRingClient, RingFabric
    from fps_b200.store.sharded_table import ShardedTable

    res = {"n_gpus": world, "keys": a.keys, "txn_per_rank": a.txn, "dim": a.dim, "lanes": a.lanes,
           "ring_capacity": a.capacity, "pull_limit": a.limit, "runs": []}
    for name, update, lock, mode in (("lockA_pull_push_add", "add", "A", "txn"),
                                     ("assign_pull_push", "assign", None, "txn"),
                                     ("assign_push_only", "assign", None, "push")):
        table = ShardedTable(a.keys, a.dim, seed=3, init_range=(0.0, 1.0))
        rings = RingFabric(table.stride, capacity=a.capacity, lanes=a.lanes)
        server = DeviceMessageServer(table, rings, update=update, lock=lock, pool_size=1 << 20)
        client = DeviceRingClient(table, rings, pull_limit=a.limit)
        if world > 1:
            dist.barrier()               # every ring exists; messages may arrive before a server polls
        server.start()
        g = torch.Generator().manual_seed(100 + rank)
        best = None
        try:
          for it in range(a.iters + 1):
            ids = torch.randint(0, a.keys, (a.txn,), generator=g).to(dev)
            d = torch.ones(a.txn, a.dim, device=dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(client.stream):
                e0.record(client.stream)
                if mode == "txn":
                    client.transact(ids, d)
                else:
                    client.push_all(ids, d)
                e1.record(client.stream)
            st = client.wait()
            ms = e0.elapsed_time(e1)
            if it > 0:
                best = ms if best is None else min(best, ms)
        except BaseException:
            server.stop_flag_only()          # never leave a persistent kernel behind a failing run
            raise
        if world > 1:
            with torch.cuda.stream(client.stream):      # every worker is done before any server stops
                dist.barrier()
            client.stream.synchronize()
        server.stop()
        if world > 1:
            t = torch.tensor([best], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = float(t.item())
        msgs = a.txn * (2 if mode == "txn" else 1)
        res["runs"].append({"case": name, "ms": best, "messages_per_s_per_gpu": msgs / best * 1e3,
                            "includes": "host-side sort of the batch by ring + the persistent client kernel",
                            "server": server.stats()})
        rings.close(); table.close()
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
