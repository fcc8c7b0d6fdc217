"""``FlinkSleepBlocker.block`` (M/utils/FlinkSleepBlocker.scala:23-36): hold a stream back for a
fixed time before its first record -- a manual ordering / back-pressure tool for model loading."""
from __future__ import annotations

import time

from ..runtime.stream import DataStream, as_stream


def block(stream, milliseconds: float) -> DataStream:
    s = as_stream(stream)

    def delayed(it):
        slept = False
        for x in it:
            if not slept:
                time.sleep(milliseconds / 1000.0)
                slept = True
            yield x

    return s._wrap(delayed)
