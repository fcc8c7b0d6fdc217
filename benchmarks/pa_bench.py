#!/usr/bin/env python
"""Passive-aggressive fused CSR kernel: examples/s and pull+push row traffic (K7)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    from fps_b200.models.pa.device import DevicePassiveAggressive

    feats, n, nnz = 5_000_000, 16384, 256
    res = {}
    from fps_b200.ops import native
    for name, L, binary in [("binary_PA", 1, True), ("binary_PA_block_kernel", 1, True), ("ova4_PAI", 4, False),
                            ("ova32_PAI", 32, False)]:
        native.lib().fps_set_pa_variant(1 if "block_kernel" in name else 0)
        pa = DevicePassiveAggressive(feats, L, binary, "PA" if binary else "PAI", 1.0)
        g = torch.Generator(device="cpu").manual_seed(0)
        cols = torch.randint(0, feats, (n * nnz,), generator=g, dtype=torch.int32).to(dev)
        vals = torch.randn(n * nnz, generator=g).to(dev)
        row_ptr = (torch.arange(n + 1, dtype=torch.int64) * nnz).to(dev)
        labels = (torch.randint(0, 2, (n,), generator=g) * 2 - 1).int().to(dev) if binary else \
            torch.randint(0, L, (n,), generator=g).int().to(dev)
        for _ in range(3):
            pa.step_csr(row_ptr, cols, vals, labels)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            pa.step_csr(row_ptr, cols, vals, labels)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        row_bytes = pa.table.stride * 4
        res[name] = {"examples_per_s": n / ms * 1e3, "ms_per_batch": ms, "nnz_per_example": nnz,
                     "pull_push_GBs": n * nnz * 2 * row_bytes / ms / 1e6, "labels": L}
        pa.check_finite(); pa.close()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
