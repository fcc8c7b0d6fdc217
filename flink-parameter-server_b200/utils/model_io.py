"""Model export / import (the reference's resume-by-reload path, SURVEY §5 "Checkpoint / resume").

Text format = what the reference's experiment mains write with ``writeAsText``:
``id:v1,v2,...`` one parameter per line (BloomFilterExp.scala:45-46; parsers in *PredictExp.scala).
Binary format = ``.npz`` with ``ids`` and ``values`` arrays (fast path for dense factor tables).
The loaded ``(id, value)`` pairs feed ``transformWithModelLoad`` / ``ShardedTable.load``.
"""
from __future__ import annotations

from typing import Callable, Iterable, Iterator, Tuple

import numpy as np


def format_param(id, value) -> str:
    if isinstance(id, tuple):
        id = ";".join(str(x) for x in id)
    if isinstance(value, (int, float)):
        return f"{id}:{value}"
    return f"{id}:{','.join(str(x) for x in (sorted(value) if isinstance(value, (set, frozenset)) else value))}"


def write_text(path: str, model: Iterable[Tuple[object, object]]) -> int:
    n = 0
    with open(path, "w") as f:
        for id, value in model:
            f.write(format_param(id, value) + "\n")
            n += 1
    return n


def read_text(path: str, id_type: Callable = int, value_type: Callable = float,
              as_set: bool = False) -> Iterator[Tuple[object, object]]:
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            k, _, v = line.partition(":")
            key = tuple(id_type(x) for x in k.split(";")) if ";" in k else id_type(k)
            vals = [value_type(x) for x in v.split(",")] if v else []
            yield key, (frozenset(vals) if as_set else (vals[0] if len(vals) == 1 and "," not in v else vals))


def write_npz(path: str, ids, values) -> None:
    np.savez(path, ids=np.asarray(ids), values=np.asarray(values))


def read_npz(path: str):
    d = np.load(path)
    return d["ids"], d["values"]
