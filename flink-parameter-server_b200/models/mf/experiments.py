"""The reference's two matrix-factorisation experiment mains (they live among its tests) as a CLI:

    python -m fps_b200.models.mf.experiments OnlineMFImplicit <input> <userVectors.out> <itemVectors.out> [backend]
    python -m fps_b200.models.mf.experiments OnlineLearnerAndTopK <input> <nDCG.csv> [backend] [periodLength]

* ``OnlineMFImplicit`` -- T/matrix/factorization/PSOnlineMatrixFactorizationImplicitTest.scala:32-98: a
  Last.fm-style implicit-feedback file (``<time> <user> <item>`` separated by blanks, every line a rating of
  1.0), online MF with 9 negative samples per event, and the two ``id;value`` vector maps (one line per
  vector component) that the reference's ``Notebooks/Tester.ipynb`` reads -- ``utils/evaluation.py`` has
  that notebook's offline precision / recall.
* ``OnlineLearnerAndTopK`` -- T/matrix/factorization/PSOnlineMatrixFactorizationAndTopKGeneratorTest.scala:27-53:
  ``<time>,<user>,<item>`` lines, online learner + a top-K recommendation per event (prequential), nDCG@K and
  hit rate per day written by the nDCG sink (``utils/metrics.py::nDCGPeriodsToCsv``).

Hyper-parameters are the constants of those mains; ``backend`` is ``local`` (default), ``native`` (first
main only) or ``device``.
"""
from __future__ import annotations

import sys
from typing import Dict, Iterable, List, Sequence

from ...api import Left, Right
from ...utils.metrics import nDCGPeriodsToCsv
from .common import Rating
from .online import psOnlineMF
from .topk import psOnlineLearnerAndGenerator

IMPLICIT = dict(numFactors=10, rangeMin=-0.1, rangeMax=0.1, learningRate=0.01, userMemory=128,
                negativeSampleRate=9, pullLimit=1500, workerParallelism=4, psParallelism=4,
                iterationWaitTime=10000)
LEARNER = dict(numFactors=10, rangeMin=-0.01, rangeMax=0.01, learningRate=0.2, userMemory=4, K=100,
               workerK=100, bucketSize=100, negativeSampleRate=9, pullLimit=800, workerParallelism=4,
               psParallelism=4, iterationWaitTime=20000)


def read_implicit(path: str) -> List[Rating]:
    """``<time> <user> <item>`` -> rating 1.0 (the main ignores the time column)."""
    out = []
    with open(path) as f:
        for line in f:
            fields = line.split()
            if len(fields) >= 3:
                out.append(Rating(int(fields[1]), int(fields[2]), 1.0, 0))
    return out


def read_week(path: str) -> List[Rating]:
    """``<time>,<user>,<item>`` -> ``Rating(user, item, 1.0, time)``."""
    out = []
    with open(path) as f:
        for line in f:
            fields = line.strip().split(",")
            if len(fields) >= 3:
                out.append(Rating(int(fields[1]), int(fields[2]), 1.0, int(fields[0])))
    return out


def write_vector_map(path: str, vectors: Dict[int, Sequence[float]]) -> None:
    """``id;value`` -- one line per component, components in order (the format ``Tester.ipynb`` parses)."""
    with open(path, "w") as f:
        for k, v in vectors.items():
            for x in v:
                f.write(f"{k};{float(x)}\n")


def last_vectors(stream: Iterable) -> (Dict[int, Sequence[float]], Dict[int, Sequence[float]]):
    """The sink of the first main: keep the LAST vector seen per user (``Left``) / item (``Right``)."""
    users, items = {}, {}
    for rec in stream:
        if isinstance(rec, Left):
            users[rec.value[0]] = rec.value[1]
        elif isinstance(rec, Right):
            items[rec.value[0]] = rec.value[1]
    return users, items


def OnlineMFImplicit(a: List[str], **overrides):
    src, user_out, item_out = a[0], a[1], a[2]
    kw = dict(IMPLICIT, **overrides)
    backend = a[3] if len(a) > 3 else "local"
    users, items = last_vectors(psOnlineMF(read_implicit(src), backend=backend, **kw))
    write_vector_map(user_out, users)
    write_vector_map(item_out, items)
    return users, items


def OnlineLearnerAndTopK(a: List[str], **overrides):
    src, out_csv = a[0], a[1]
    backend = a[2] if len(a) > 2 else "local"
    period = int(a[3]) if len(a) > 3 else 86400
    kw = dict(LEARNER, **overrides)
    topk = psOnlineLearnerAndGenerator(read_week(src), backend=backend, **kw)
    return nDCGPeriodsToCsv(topk, out_csv, period)


MAINS = {"OnlineMFImplicit": OnlineMFImplicit, "OnlineLearnerAndTopK": OnlineLearnerAndTopK,
         # the reference's object names
         "PSOnlineMatrixFactorizationImplicitTest": OnlineMFImplicit,
         "PSOnlineMatrixFactorizationAndTopKGeneratorTest": OnlineLearnerAndTopK}


def main(argv: List[str]) -> int:
    if len(argv) < 2 or argv[0] not in MAINS:
        print(__doc__)
        return 2
    MAINS[argv[0]](argv[1:])
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
