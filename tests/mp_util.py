"""Helpers for the multi-rank device checks (``tests/mp_*_check.py``).

On a box with at least ``WORLD_SIZE`` GPUs every rank takes its own GPU and the control plane is NCCL.
With fewer GPUs (the driver's single-GPU test box), or with ``FPS_SHARE_GPU=1``, the ranks SHARE the
visible GPUs: the control plane is gloo, and the data plane is unchanged -- CUDA IPC maps the other
processes' shards exactly as it maps peer GPUs, so the one-sided kernels, rings and the replica
exchange run the same code (peer traffic then stays inside one HBM instead of crossing NVLink).
"""
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def init_dist():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    n_dev = torch.cuda.device_count()
    shared = os.environ.get("FPS_SHARE_GPU") == "1" or n_dev < world
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if shared:
        dev = torch.device("cuda", local % n_dev)
        torch.cuda.set_device(dev)
        dist.init_process_group("gloo")
    else:
        dev = torch.device("cuda", local)
        torch.cuda.set_device(dev)
        dist.init_process_group("nccl", device_id=dev)
    return rank, world, dev, shared


def all_reduce_sum(t: torch.Tensor) -> torch.Tensor:
    """Sum over ranks, backend-agnostic (gloo reduces a host copy)."""
    if dist.get_backend() == "nccl":
        dist.all_reduce(t)
        return t
    h = t.detach().cpu()
    dist.all_reduce(h)
    t.copy_(h)
    return t


def all_gather_cat(t: torch.Tensor) -> torch.Tensor:
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, t.detach().cpu())
    return torch.cat([p.to(t.device) for p in parts])


def launch_cmd(script: str, world: int, port: int):
    return [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
            "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(REPO, "tests", script)]
