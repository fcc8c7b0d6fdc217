"""A small ``DataStream``-shaped iterator abstraction (the slice of Flink's L0 the PS needs).

A :class:`DataStream` is a set of parallel *sources* (iterables, possibly unbounded) plus, per
source, a *routing* rule saying how its records reach the subtasks of the consuming operator:

============================  ==============================================================
``forward``                   source ``i`` -> subtask ``i % n``                (Flink forward)
``rebalance``                 round robin                                     (``rebalance``)
``broadcast``                 every record to every subtask                   (``broadcast``)
``partition_custom(p, key)``  subtask ``p(key(record), n)``                   (``partitionCustom``)
============================  ==============================================================

``map / flat_map / filter`` are applied lazily inside the source iterators.  A source may belong
to an *EOF group*: when every source of the group is exhausted the engine delivers the group's
marker record to **every** consuming subtask after all of the group's records -- this is the
capability of ``FlinkEOF.flatMapWithEOF`` (M/utils/FlinkEOF.scala:71-122) and of the model-load
EOF fan-out (FPS:744-762) without per-upstream-subtask EOF counting.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field
from typing import Any, Callable, Iterable, Iterator, List, Optional, Sequence

_group_ids = itertools.count(1)


@dataclass
class Routing:
    kind: str = "rebalance"                       # forward | rebalance | broadcast | custom
    partitioner: Optional[Callable[[Any, int], int]] = None
    key: Optional[Callable[[Any], Any]] = None

    def targets(self, record: Any, n: int, source_index: int, counter: List[int]) -> Sequence[int]:
        if self.kind == "forward":
            return (source_index % n,)
        if self.kind == "rebalance":
            t = counter[0] % n
            counter[0] += 1
            return (t,)
        if self.kind == "broadcast":
            return range(n)
        if self.kind == "custom":
            k = self.key(record) if self.key is not None else record
            t = int(self.partitioner(k, n))
            if not 0 <= t < n:
                raise RuntimeError(f"partitioner returned {t} for {n} partitions")
            return (t,)
        raise ValueError(self.kind)


@dataclass
class Source:
    make_iter: Callable[[], Iterator[Any]]
    routing: Routing = field(default_factory=Routing)
    group: Optional[int] = None                   # EOF group id
    index: int = 0


class DataStream:
    def __init__(self, sources: List[Source], eof_markers: Optional[dict] = None):
        self.sources = sources
        self.eof_markers = dict(eof_markers or {})   # group id -> marker record factory

    # ---- construction ---------------------------------------------------------------
    @staticmethod
    def from_collection(items: Iterable[Any]) -> "DataStream":
        items = items if isinstance(items, (list, tuple)) else list(items)
        return DataStream([Source(lambda: iter(items), Routing("rebalance"), index=0)])

    @staticmethod
    def from_iterable(it: Iterable[Any]) -> "DataStream":
        """One (possibly unbounded, single-pass) source."""
        return DataStream([Source(lambda: iter(it), Routing("rebalance"), index=0)])

    @staticmethod
    def from_parallel(parts: Sequence[Iterable[Any]]) -> "DataStream":
        """``len(parts)`` parallel sources, forwarded source i -> subtask i."""
        return DataStream([Source((lambda p=p: iter(p)), Routing("forward"), index=i)
                           for i, p in enumerate(parts)])

    # ---- record-wise lazy operators ------------------------------------------------------
    def _wrap(self, f: Callable[[Iterator[Any]], Iterator[Any]]) -> "DataStream":
        new = [Source((lambda s=s: f(s.make_iter())), s.routing, s.group, s.index)
               for s in self.sources]
        return DataStream(new, self.eof_markers)

    def map(self, fn: Callable[[Any], Any]) -> "DataStream":
        return self._wrap(lambda it: (fn(x) for x in it))

    def flat_map(self, fn: Callable[[Any], Iterable[Any]]) -> "DataStream":
        return self._wrap(lambda it: (y for x in it for y in fn(x)))

    flatMap = flat_map

    def filter(self, fn: Callable[[Any], bool]) -> "DataStream":
        return self._wrap(lambda it: (x for x in it if fn(x)))

    # ---- routing ------------------------------------------------------------------------
    def _route(self, r: Routing) -> "DataStream":
        return DataStream([Source(s.make_iter, r, s.group, s.index) for s in self.sources],
                          self.eof_markers)

    def forward(self) -> "DataStream":
        return self._route(Routing("forward"))

    def rebalance(self) -> "DataStream":
        return self._route(Routing("rebalance"))

    def broadcast(self) -> "DataStream":
        return self._route(Routing("broadcast"))

    def partition_custom(self, partitioner: Callable[[Any, int], int],
                         key: Optional[Callable[[Any], Any]] = None) -> "DataStream":
        return self._route(Routing("custom", partitioner, key))

    partitionCustom = partition_custom

    # ---- EOF / union --------------------------------------------------------------------
    def with_eof(self, marker: Callable[[], Any]) -> "DataStream":
        """After ALL sources of this stream end, deliver ``marker()`` to every consumer subtask."""
        gid = next(_group_ids)
        new = [Source(s.make_iter, s.routing, gid, s.index) for s in self.sources]
        m = dict(self.eof_markers)
        m[gid] = marker
        return DataStream(new, m)

    def union(self, *others: "DataStream") -> "DataStream":
        srcs = list(self.sources)
        marks = dict(self.eof_markers)
        for o in others:
            srcs.extend(o.sources)
            marks.update(o.eof_markers)
        srcs = [Source(s.make_iter, s.routing, s.group, i) if s.routing.kind != "forward" else s
                for i, s in enumerate(srcs)]
        return DataStream(srcs, marks)

    # ---- eager helpers ------------------------------------------------------------------
    def collect(self) -> List[Any]:
        out: List[Any] = []
        for s in self.sources:
            out.extend(s.make_iter())
        return out

    def __iter__(self) -> Iterator[Any]:
        return iter(self.collect())


def as_stream(data: Any) -> DataStream:
    if isinstance(data, DataStream):
        return data
    if isinstance(data, (list, tuple)):
        return DataStream.from_collection(data)
    return DataStream.from_iterable(data)


class ResultStream(DataStream):
    """The ``DataStream[Either[WOut, PSOut]]`` returned by ``transform``: materialised results."""

    def __init__(self, results: List[Any]):
        super().__init__([Source(lambda: iter(results), Routing("rebalance"))])
        self.results = results

    def collect(self) -> List[Any]:
        return list(self.results)

    def worker_outputs(self) -> List[Any]:
        return [r.value for r in self.results if r.is_left]

    def ps_outputs(self) -> List[Any]:
        return [r.value for r in self.results if r.is_right]

    def __len__(self) -> int:
        return len(self.results)
