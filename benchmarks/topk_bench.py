#!/usr/bin/env python
"""Top-K scoring benchmark (K6): B queries pulled from the PS x N local items, k=64, K=100.
Reports device time per batch, TF32 tensor throughput of the two GEMM passes and the comparison with
torch (cuBLAS fp32 matmul + torch.topk) on the same data."""
import argparse
import json
import os

import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def t_ms(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--items", type=int, default=1_000_000)
    ap.add_argument("--users", type=int, default=1_000_000)
    ap.add_argument("--queries", type=int, default=2048)
    ap.add_argument("--factors", type=int, default=64)
    ap.add_argument("--K", type=int, default=100)
    ap.add_argument("--skew", type=float, default=0.0,
                    help="> 0: log-normal item lengths with this sigma (popularity skew) instead of U[0.05, 2.05)")
    ap.add_argument("--pass1-fraction", type=float, default=0.0,
                    help="> 0: also time DeviceTopK(pass1_fraction=f) (theta from a prefix of the tiles)")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    from fps_b200.models.mf.device_topk import DeviceTopK
    from fps_b200.ops import native
    from fps_b200.store.sharded_table import ShardedTable

    users = ShardedTable(a.users, a.factors, seed=1, init_range=(-1, 1))
    items = ShardedTable(a.items, a.factors, seed=2, init_range=(-1, 1))
    # LEMP-style: make item lengths heterogeneous so tile pruning has something to prune
    if a.skew > 0:
        items.local.mul_(torch.exp(torch.randn(items.local.shape[0], 1, device=dev) * a.skew))
    else:
        items.local.mul_(torch.rand(items.local.shape[0], 1, device=dev) * 2 + 0.05)
    tk = DeviceTopK(items.local)
    q = torch.randint(0, a.users, (a.queries,), device=dev)
    tile_max = torch.empty((a.queries, tk.n_tiles), dtype=torch.float32, device=dev)
    ms_pass1 = t_ms(lambda: native.topk_mma(items.local, 1, q_ids=q, q_tab=users.table_c, tile_max=tile_max))
    ms_total = t_ms(lambda: tk.topk(a.K, q_ids=q, q_table=users), iters=5, warm=2)

    def torch_ref():
        u = users.pull(q)
        return torch.topk(u @ items.local[:, : a.factors].T, a.K, dim=1)

    ms_torch = t_ms(torch_ref, iters=5, warm=2)
    tkp = DeviceTopK(items.local, sort_by_length=True)       # LEMP LENGTH bound at tile granularity
    ms_pruned = t_ms(lambda: tkp.topk(a.K, q_ids=q, q_table=users), iters=5, warm=2)
    p1, p2 = tkp.last_tiles_scored
    ms_frac = None
    if a.pass1_fraction > 0:
        tkf = DeviceTopK(items.local, pass1_fraction=a.pass1_fraction)
        ms_frac = t_ms(lambda: tkf.topk(a.K, q_ids=q, q_table=users), iters=5, warm=2)
        ref_s, _ = tk.topk(a.K, q_ids=q, q_table=users)
        got_s, _ = tkf.topk(a.K, q_ids=q, q_table=users)
        assert torch.equal(ref_s, got_s), "pass1_fraction changed the result"
    for name, t in (("plain", tk), ("pruned", tkp)):       # per-stage breakdown (synchronising trace)
        t.trace = []; t._t0 = None
        t.topk(a.K, q_ids=q, q_table=users)
        print(name, " ".join(f"{lbl}={ms:.3f}" for lbl, ms in t.trace), file=sys.stderr)
        t.trace = None
    flops = 2.0 * a.queries * a.items * a.factors
    print(json.dumps({"queries": a.queries, "items": a.items, "factors": a.factors, "K": a.K,
                      "pass1_ms": ms_pass1, "pass1_tf32_TFLOPs": flops / ms_pass1 / 1e9,
                      "topk_total_ms": ms_total, "queries_per_s": a.queries / ms_total * 1e3,
                      "pass1_fraction": a.pass1_fraction, "pass1_fraction_total_ms": ms_frac,
                      "length_pruned_total_ms": ms_pruned, "tiles": tkp.n_tiles, "tiles_pass1": p1,
                      "tiles_pass2": p2, "skew": a.skew,
                      "torch_matmul_topk_ms": ms_torch, "speedup_vs_torch": ms_torch / ms_total}))


if __name__ == "__main__":
    main()
