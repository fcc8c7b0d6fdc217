#!/usr/bin/env python
"""Launch-bound streaming regime: small micro-batches, eager launches vs one CUDA-graph replay per step
(`DeviceOnlineMF.make_graph_step`).  Reports us/step and updates/s for both."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    from fps_b200.models.mf.device import DeviceOnlineMF
    from fps_b200.ops import native

    m = DeviceOnlineMF(1_000_000, 100_000, 64, learning_rate=0.01, seed=1)
    res = {}
    g = torch.Generator().manual_seed(0)
    for bs in (1024, 4096, 16384, 65536):
        u = torch.randint(0, 1_000_000, (bs,), generator=g, dtype=torch.int32)
        i = torch.randint(0, 100_000, (bs,), generator=g, dtype=torch.int32)
        rec = native.pack_ratings(u.to(dev), i.to(dev), torch.rand(bs, generator=g).to(dev))
        iters = 500

        def eager():
            for _ in range(iters):
                m.step(rec)

        (static,), replay = m.make_graph_step(bs, packed=True)
        static.copy_(rec)

        def graphed():
            for _ in range(iters):
                replay()

        out = {}
        for name, fn in (("eager", eager), ("graph", graphed)):
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / iters
            out[name] = {"us_per_step": us, "updates_per_s": bs / us * 1e6}
        out["speedup"] = out["eager"]["us_per_step"] / out["graph"]["us_per_step"]
        res[str(bs)] = out
    m.check_finite()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
