"""Rate-replay file source (capability of M/matrix/factorization/utils/InputSource.scala:11-68).

Replays timestamped events at ``servingSpeed`` x real time: event ``e`` is emitted at
``servingStart + (e.time - dataStart) / servingSpeed``; after ``simulationEndTime`` (if given)
events are replayed at 1x.  ``cancel()`` works (the reference's is ``???``).
"""
from __future__ import annotations

import threading
import time
from typing import Any, Callable, Iterable, Iterator, List, Optional


class EventWithTimestamp:
    def getEventTime(self) -> int:
        raise NotImplementedError


class InputSource:
    def __init__(self, dataFilePath: Optional[str], servingSpeed: float,
                 fromString: Optional[Callable[[str], Any]] = None,
                 baseDataStartTime: Optional[int] = None, simulationEndTime: Optional[int] = None,
                 events: Optional[Iterable[Any]] = None,
                 event_time: Optional[Callable[[Any], int]] = None,
                 clock: Callable[[], float] = time.time, sleep: Callable[[float], None] = time.sleep):
        self.path, self.speed, self.fromString = dataFilePath, float(servingSpeed), fromString
        self.baseDataStartTime = baseDataStartTime
        self.simEndTime = simulationEndTime or 0
        self._events = events
        self._event_time = event_time or (lambda e: e.getEventTime())
        self._clock, self._sleep = clock, sleep
        self._cancel = threading.Event()

    def cancel(self) -> None:
        self._cancel.set()

    def _load(self) -> List[Any]:
        if self._events is not None:
            return list(self._events)
        with open(self.path) as f:
            return [self.fromString(line.rstrip("\n")) for line in f if line.strip()]

    def _serving_time(self, start_ms: float, data_start: int, t: int) -> float:
        if self.simEndTime != 0 and t >= self.simEndTime:
            return (start_ms + self.simEndTime / self.speed) + (t - self.simEndTime) - data_start / self.speed
        return start_ms + (t - data_start) / self.speed

    def __iter__(self) -> Iterator[Any]:
        events = self._load()
        if not events:
            return
        data_start = self.baseDataStartTime if self.baseDataStartTime is not None \
            else self._event_time(events[0])
        start_ms = self._clock() * 1000.0
        for e in events:
            if self._cancel.is_set():
                return
            wait = self._serving_time(start_ms, data_start, self._event_time(e)) - self._clock() * 1000.0
            if wait > 0:
                self._sleep(wait / 1000.0)
            yield e
