"""Wire messages (M/entities/Messages.scala:3-8).

On the device tier there are no messages at all (a pull is a peer load, a push is a peer
reduction); these records are the wire format of the generic host tier and of the CPU backend.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Union

from ..api import Left, Right


@dataclass(frozen=True)
class Pull:
    paramId: Any


@dataclass(frozen=True)
class Push:
    paramId: Any
    delta: Any


@dataclass(frozen=True)
class PullAnswer:
    paramId: Any
    param: Any


@dataclass(frozen=True)
class WorkerToPS:
    """``msg`` is ``Left(Pull)`` or ``Right(Push)``; ``workerPartitionIndex`` routes the answer."""

    workerPartitionIndex: int
    msg: Union[Left, Right]

    @property
    def paramId(self):
        return self.msg.value.paramId


@dataclass(frozen=True)
class PSToWorker:
    workerPartitionIndex: int
    msg: PullAnswer
