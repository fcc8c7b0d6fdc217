"""Host-side placement for the ingest path (E1 of SURVEY §2.11: source -> worker).

A worker process feeds its GPU from pinned host memory; on a two-socket box the copy runs at PCIe speed
only if those pages live on the socket the GPU hangs off -- otherwise every byte crosses the inter-socket
link first, and with eight workers the link, not PCIe, bounds the end-to-end rate.  ``torchrun`` does not
bind its children, so a launcher (or ``bench.py``) calls :func:`bind_to_gpu_node` once per process,
BEFORE it allocates pinned buffers: the calling thread (and every thread it starts later) is restricted to
the CPUs of the GPU's NUMA node and its memory policy prefers that node, which is what
``numactl --cpunodebind=N --preferred=N`` would do.

Everything here is best effort: missing sysfs entries, a single-node box, a cpuset that excludes the
node's CPUs or a refused syscall leave the process exactly as it was (the returned dict says what
happened).  Nothing in the reference corresponds to this (Flink schedules its own task slots).
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional, Set

_SYS_PCI = "/sys/bus/pci/devices"
_SYS_NODE = "/sys/devices/system/node"
_MPOL_PREFERRED = 1
_SYS_set_mempolicy = 238          # x86_64


def parse_cpulist(text: str) -> Set[int]:
    """``"0-3,8,10-11"`` -> ``{0, 1, 2, 3, 8, 10, 11}`` (the sysfs ``cpulist`` format)."""
    cpus: Set[int] = set()
    for part in text.strip().split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            lo, hi = part.split("-", 1)
            cpus.update(range(int(lo), int(hi) + 1))
        else:
            cpus.add(int(part))
    return cpus


def pci_address(device_index: int) -> Optional[str]:
    """``domain:bus:device.0`` of a visible CUDA device, in sysfs spelling."""
    try:
        import torch

        p = torch.cuda.get_device_properties(device_index)
        return "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
    except Exception:
        return None


def gpu_numa_node(device_index: int, pci_root: str = _SYS_PCI) -> Optional[int]:
    """NUMA node the GPU is attached to, ``None`` if unknown (VMs report -1)."""
    addr = pci_address(device_index)
    if addr is None:
        return None
    try:
        with open(os.path.join(pci_root, addr, "numa_node")) as f:
            node = int(f.read().strip())
    except (OSError, ValueError):
        return None
    return node if node >= 0 else None


def node_cpus(node: int, node_root: str = _SYS_NODE) -> Set[int]:
    try:
        with open(os.path.join(node_root, f"node{int(node)}", "cpulist")) as f:
            return parse_cpulist(f.read())
    except (OSError, ValueError):
        return set()


def _prefer_node(node: int) -> bool:
    """``set_mempolicy(MPOL_PREFERRED, {node})`` for the calling thread; False if refused."""
    if not 0 <= node < 128:
        return False
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        libc.syscall.restype = ctypes.c_long
        mask = (ctypes.c_ulong * 2)(0, 0)
        mask[node // 64] = 1 << (node % 64)
        rc = libc.syscall(ctypes.c_long(_SYS_set_mempolicy), ctypes.c_int(_MPOL_PREFERRED),
                          ctypes.byref(mask), ctypes.c_ulong(129))
        return rc == 0
    except Exception:
        return False


def bind_to_node(node: Optional[int], min_cpus: int = 2, node_root: str = _SYS_NODE,
                 set_policy: bool = True) -> Dict[str, object]:
    """Restrict the calling thread to ``node``'s CPUs (intersected with what it may use now) and prefer
    the node's memory.  Returns ``{"node", "cpus", "bound", "mempolicy", "why"}``."""
    info: Dict[str, object] = {"node": node, "cpus": None, "bound": False, "mempolicy": False, "why": None}
    if node is None:
        info["why"] = "NUMA node of the GPU unknown"
        return info
    try:
        n_nodes = len([d for d in os.listdir(node_root) if d.startswith("node") and d[4:].isdigit()])
    except OSError:
        n_nodes = 0
    if n_nodes <= 1:
        info["why"] = "single NUMA node"
        return info
    try:
        allowed = os.sched_getaffinity(0)
    except (AttributeError, OSError):
        info["why"] = "sched_getaffinity unavailable"
        return info
    cpus = node_cpus(node, node_root) & allowed
    info["cpus"] = len(cpus)
    if len(cpus) >= max(1, int(min_cpus)):
        try:
            os.sched_setaffinity(0, cpus)
            info["bound"] = True
        except OSError as e:
            info["why"] = f"sched_setaffinity: {e}"
    else:
        info["why"] = "the cpuset leaves fewer than %d CPUs of node %d" % (min_cpus, node)
    if set_policy:
        info["mempolicy"] = _prefer_node(int(node))
    return info


def bind_to_gpu_node(device_index: int, min_cpus: int = 2) -> Dict[str, object]:
    """Place the calling thread next to CUDA device ``device_index`` (see the module docstring).  Honour
    ``FPS_NUMA_BIND=0`` (leave the placement to the launcher, e.g. ``numactl``)."""
    if os.environ.get("FPS_NUMA_BIND", "1") == "0":
        return {"node": None, "cpus": None, "bound": False, "mempolicy": False, "why": "FPS_NUMA_BIND=0"}
    return bind_to_node(gpu_numa_node(device_index), min_cpus=min_cpus)
