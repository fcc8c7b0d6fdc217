"""Shared helper: gather a (scores, ids) pair of per-rank partial top-K lists with :class:`P2PGather`."""
from __future__ import annotations

import torch


def gather_pair(owner, sc: torch.Tensor, gids: torch.Tensor, dst, group, device):
    """One message per rank: fp32 score bits followed by the int64 ids (as int32 pairs).  The gather
    object lives on ``owner`` (re-created, collectively, when a larger message arrives).  Returns
    ``([scores per rank], [ids per rank])`` on the receiving rank(s), ``None`` elsewhere."""
    from .parallel.fabric import P2PGather

    n_sc = sc.numel()
    n_pad = n_sc + (n_sc & 1)            # keep the int64 ids 8-byte aligned inside the message
    head = sc.contiguous().view(torch.int32).reshape(-1)
    if n_pad != n_sc:
        head = torch.nn.functional.pad(head, (0, 1))
    msg = torch.cat([head, gids.contiguous().to(torch.int64).view(torch.int32).reshape(-1)])
    need = msg.numel() * 4
    g = getattr(owner, "_p2p_gather", None)
    if g is None or g.slot < need:
        if g is not None:
            g.close()
        g = P2PGather(max(need, 1 << 16), group=group, device=device)
        owner._p2p_gather = g
    parts = g.gather(msg, dst=dst)
    if parts is None:
        return None
    scs = [p[:n_sc].view(torch.float32).reshape(sc.shape) for p in parts]
    ids = [p[n_pad:].view(torch.int64).reshape(gids.shape) for p in parts]
    return scs, ids
