"""End-of-input detection for finite streams (capability of M/utils/FlinkEOF.scala).

``flatMapWithEOF(stream, fn, downstreamParallelism, partitioner, key)`` applies ``fn`` on
``downstreamParallelism`` parallel instances, each of which gets ``onEOF(collector)`` called exactly
once after *all* records of *all* upstream subtasks have been seen.  Unlike the reference an empty
upstream subtask is fine (FlinkEOF.scala throws ``UnsupportedOperationException``, :97-100).
"""
from __future__ import annotations

import copy
from typing import Any, Callable, List, Optional

from ..runtime.stream import DataStream, Routing, Source, as_stream


class EOF:
    """End-of-stream marker delivered to consumers (``Left(EOF)`` in psOfflineMF)."""

    __slots__ = ()

    def __repr__(self):
        return "EOF"

    def __eq__(self, other):
        return isinstance(other, EOF)

    def __hash__(self):
        return hash("fps.EOF")


class EOFHandler:
    """Mixin for flat-map functions: ``flatMap(value, collect)`` + ``onEOF(collect)``."""

    def flatMap(self, value: Any, collect: Callable[[Any], None]) -> None:
        raise NotImplementedError

    def onEOF(self, collect: Callable[[Any], None]) -> None:
        raise NotImplementedError


def flatMapWithEOF(stream, flatMapFunction: EOFHandler, downstreamParallelism: int,
                   partitioner: Callable[[Any, int], int],
                   partitionerFunc: Callable[[Any], Any]) -> DataStream:
    """Partition ``stream`` into ``downstreamParallelism`` instances of ``flatMapFunction``; each
    instance emits its ``onEOF`` output after its last record.  Result: ``downstreamParallelism``
    forward-routed parallel sources (so a following PS ``transform`` keeps the partitioning)."""
    src = as_stream(stream)
    n = int(downstreamParallelism)
    state = {"parts": None}

    def materialise():
        if state["parts"] is None:
            fns = [copy.deepcopy(flatMapFunction) for _ in range(n)]
            outs: List[List[Any]] = [[] for _ in range(n)]
            for rec in src.collect():
                t = int(partitioner(partitionerFunc(rec), n))
                fns[t].flatMap(rec, outs[t].append)
            for t in range(n):
                fns[t].onEOF(outs[t].append)
            state["parts"] = outs
        return state["parts"]

    return DataStream([Source((lambda i=i: iter(materialise()[i])), Routing("forward"), index=i)
                       for i in range(n)])


def with_eof(stream, partitioner: Optional[Callable[[Any, int], int]] = None,
             key: Optional[Callable[[Any], Any]] = None) -> DataStream:
    """Streaming form: route records (custom partitioner or rebalance) and deliver one ``EOF()``
    to every consumer subtask after the last record of the whole stream."""
    s = as_stream(stream)
    s = s.partition_custom(partitioner, key) if partitioner is not None else s
    return s.with_eof(EOF)


flat_map_with_eof = flatMapWithEOF
