"""Convergence of the replica mode vs its staleness knob (run under torchrun; ranks may share one GPU:
``FPS_SHARE_GPU=1`` -- the schedule of exchanges is step based, so quality does not depend on timing).

    FPS_SHARE_GPU=1 python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 \
        benchmarks/quality_sweep.py --sync 1,2,4,8 --updates-per-user 400

Prints one JSON line: held-out RMSE of ONE worker alone, N workers in direct one-sided mode and N workers
in replica mode for every ``sync_every``, all on the same synthetic low-rank stream and update budget.
"""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402


def main():
    import torch
    import torch.distributed as dist

    argv = sys.argv[1:]
    sync = [1, 2, 4, 8]
    if "--sync" in argv:
        i = argv.index("--sync"); sync = [int(x) for x in argv[i + 1].split(",")]; del argv[i:i + 2]
    upu = [400.0]
    if "--updates-per-user" in argv:       # comma list = checkpoints of one run per mode
        i = argv.index("--updates-per-user"); upu = [float(x) for x in argv[i + 1].split(",")]; del argv[i:i + 2]
    skip_direct = "--skip-direct" in argv
    if skip_direct:
        argv.remove("--skip-direct")
    sys.argv = [sys.argv[0]] + argv
    a = bench.parse()
    a.quality_updates_per_user = max(upu)
    a.skip_direct_quality = skip_direct
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n_dev = torch.cuda.device_count()
    shared = world > 1 and (os.environ.get("FPS_SHARE_GPU") == "1" or n_dev < world)
    dev = torch.device("cuda", local % n_dev if shared else local)
    torch.cuda.set_device(dev)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if world > 1:
        dist.init_process_group("gloo" if shared else "nccl", **({} if shared else {"device_id": dev}))
    from fps_b200.models.mf.device import DeviceOnlineMF, ERR_PLAIN

    out = bench.quality_gate(a, world, rank, dev, shared, DeviceOnlineMF, ERR_PLAIN, sync_list=sync,
                             checkpoints=upu)
    if rank == 0:
        out.update(n_workers=world, users=a.users, items=a.items, factors=a.factors, batch=a.batch, shared_gpu=shared)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
