#!/usr/bin/env python
"""BASELINE.json configs 3 and 4 on one or more GPUs (torchrun for N > 1):
  * word2vec skip-gram negative sampling, dim=300 (push fused with paramUpdate)   -> pairs/s
  * wide-&-deep CTR, embedding table on the PS, pull-limiter=64                    -> examples/s
"""
import os

import argparse
import json
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--vocab", type=int, default=1_000_000)
    ap.add_argument("--pairs", type=int, default=1 << 20)
    ap.add_argument("--slots", type=int, default=100_000_000)
    ap.add_argument("--ctr-batch", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=30)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    lr_ = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr_); dev = torch.device("cuda", lr_)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from fps_b200.models.ctr import DeviceWideAndDeep
    from fps_b200.models.w2v import DeviceSkipGram

    def timed(fn, steps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    g = torch.Generator(device="cpu").manual_seed(rank)
    m = DeviceSkipGram(a.vocab, 300, learning_rate=0.025, negative=5, seed=1)
    c = torch.randint(0, a.vocab, (a.pairs,), generator=g, dtype=torch.int32).to(dev)
    o = torch.randint(0, a.vocab, (a.pairs,), generator=g, dtype=torch.int32).to(dev)
    ms = timed(lambda: m.step(c, o), a.steps)
    m.flush(); torch.cuda.synchronize()
    m.check_finite()
    upd = a.pairs * 6 * world
    # per update: 2 rows of 1200 B pulled + 2 rows pushed
    res = {"w2v": {"dim": 300, "vocab": a.vocab, "negative": 5, "n_gpus": world, "ms_per_step": ms,
                   "updates_per_s": upd / ms * 1e3, "replica_cache": m.rep_in is not None, "row_GBs_per_gpu": a.pairs * 6 * 4 * 1200 / ms / 1e6}}
    m.close()
    ctr = DeviceWideAndDeep(a.slots, 26, emb_dim=8, hidden=256, learning_rate=0.05, pull_limit=64, seed=1)
    ids = torch.randint(0, a.slots, (a.ctr_batch, 26), generator=g).to(dev)
    y = (torch.rand(a.ctr_batch, generator=g) < 0.3).float().to(dev)
    ms = timed(lambda: ctr.step(ids, y), a.steps)
    res["ctr"] = {"slots": a.slots, "fields": 26, "emb_dim": 8, "pull_limit": 64, "batch": a.ctr_batch,
                  "limiter": "device credit counter inside the gather kernel", "credit_stalls": ctr.credit_stalls(),
                  "tower": "fused hand-written kernel (fps_ctr.cu), loss kept on the device",
                  "n_gpus": world, "ms_per_step": ms, "examples_per_s": a.ctr_batch * world / ms * 1e3}
    ctr.close()
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
