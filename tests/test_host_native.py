"""Native host runtime (libfps_host.so): partitioner/packer and key interner."""
import numpy as np
import torch

from fps_b200.ops import host, native


def test_partition_pack_matches_torch_reference():
    g = torch.Generator().manual_seed(0)
    n, W = 100_003, 4
    u = torch.randint(0, 1 << 20, (n,), generator=g, dtype=torch.int32)
    i = torch.randint(0, 1 << 18, (n,), generator=g, dtype=torch.int32)
    r = torch.rand(n, generator=g) * 5
    parts = host.partition_pack(u, i, r, W, pin=False, threads=3)
    ref = native.pack_ratings(u, i, r)
    for w in range(W):
        sel = (u % W) == w
        assert torch.equal(parts[w], ref[sel])           # stable order, identical fp16 rounding
    assert sum(p.numel() for p in parts) == n


def test_native_interner():
    it = host.NativeInterner()
    a = it.map([10, -5, 10, 7])
    assert a.tolist() == [0, 1, 0, 2] and len(it) == 3
    assert it.map([7, 99], insert=False).tolist() == [2, -1]
    assert it.keys().tolist() == [10, -5, 7]
