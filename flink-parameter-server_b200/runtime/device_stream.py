"""Host -> device micro-batch pipeline (edge E1 of SURVEY §2.11: data ingest).

The reference feeds workers record-by-record through Flink's network stack; here a worker is fed
micro-batches: pinned host tensors are copied on a dedicated copy stream, double buffered, so the
H2D transfer of batch i+1 overlaps the fused kernel of batch i on the worker's compute stream.
"""
from __future__ import annotations

from typing import Iterable, Iterator, List, Sequence, Tuple

import torch


def pin(t: torch.Tensor) -> torch.Tensor:
    if t.is_cuda:
        raise ValueError("expected a host tensor")
    return t if t.is_pinned() else t.contiguous().pin_memory()


class DevicePrefetcher:
    """Iterate device copies of host batches with ``depth`` batches in flight."""

    def __init__(self, batches: Iterable[Sequence[torch.Tensor]], device: torch.device,
                 depth: int = 2):
        self.it = iter(batches)
        self.device = device
        self.depth = max(1, depth)
        self.copy_stream = torch.cuda.Stream(device=device)
        self.h2d_bytes = 0
        self._queue: List[Tuple[Tuple[torch.Tensor, ...], torch.cuda.Event]] = []
        self._slots: List[List[torch.Tensor]] = []
        self._slot_free: List[torch.cuda.Event] = []
        self._next_slot = 0

    def _stage(self) -> bool:
        try:
            host = next(self.it)
        except StopIteration:
            return False
        slot = self._next_slot % (self.depth + 1)
        self._next_slot += 1
        if slot >= len(self._slots):
            self._slots.append([torch.empty_like(h, device=self.device) for h in host])
            self._slot_free.append(None)
        bufs = self._slots[slot]
        if any(b.shape != h.shape or b.dtype != h.dtype for b, h in zip(bufs, host)):
            bufs = [torch.empty_like(h, device=self.device) for h in host]
            self._slots[slot] = bufs
        with torch.cuda.stream(self.copy_stream):
            if self._slot_free[slot] is not None:
                self.copy_stream.wait_event(self._slot_free[slot])  # consumer done with slot
            for b, h in zip(bufs, host):
                b.copy_(h, non_blocking=True)
                self.h2d_bytes += h.numel() * h.element_size()
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self._queue.append((tuple(bufs), ev, slot))
        return True

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, ...]]:
        for _ in range(self.depth):
            if not self._stage():
                break
        while self._queue:
            bufs, ev, slot = self._queue.pop(0)
            torch.cuda.current_stream(self.device).wait_event(ev)
            yield bufs
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(self.device))
            self._slot_free[slot] = done
            self._stage()
