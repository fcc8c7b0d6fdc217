"""Quality metrics + lightweight counters (the reference only has these inside its tests:
T/matrix/factorization/sink/nDCGSink.scala:192-272, T/test/utils/PassiveAggressive*ModelEvaluation.scala)."""
from __future__ import annotations

import math
import threading
from collections import defaultdict
from typing import Dict, Iterable, List, Sequence, Tuple


def rmse(pairs: Iterable[Tuple[float, float]]) -> float:
    se, n = 0.0, 0
    for pred, target in pairs:
        se += (pred - target) ** 2
        n += 1
    return math.sqrt(se / max(n, 1))


def accuracy(pairs: Iterable[Tuple[object, object]]) -> float:
    ok, n = 0, 0
    for pred, target in pairs:
        ok += int(pred == target)
        n += 1
    return ok / max(n, 1)


def dcg_at_k(ranked_items: Sequence[int], relevant_item: int, k: int) -> float:
    """Single relevant item (the next item the user interacts with), like the nDCG sink."""
    for rank, it in enumerate(ranked_items[:k]):
        if it == relevant_item:
            return 1.0 / math.log2(rank + 2)
    return 0.0


class NDCGSink:
    """Prequential evaluation of top-K lists: every ``(user, item, time, topK)`` record scores the list
    that was produced BEFORE the model saw the rating; aggregates nDCG@K and hit-rate per period
    (``period_of(time)``, e.g. days)."""

    def __init__(self, K: int, period_of=lambda t: 0):
        self.K, self.period_of = K, period_of
        self.sum_ndcg: Dict[int, float] = defaultdict(float)
        self.hits: Dict[int, int] = defaultdict(int)
        self.count: Dict[int, int] = defaultdict(int)

    def invoke(self, record) -> None:
        _user, item, ts, topk = record
        p = self.period_of(ts)
        items = [i for _, i in topk]
        g = dcg_at_k(items, item, self.K)
        self.sum_ndcg[p] += g
        self.hits[p] += int(g > 0)
        self.count[p] += 1

    def result(self) -> List[Tuple[int, float, float, int]]:
        """``[(period, nDCG@K, hitRate, events)]`` sorted by period."""
        return [(p, self.sum_ndcg[p] / self.count[p], self.hits[p] / self.count[p], self.count[p])
                for p in sorted(self.count)]

    def to_csv(self, path: str) -> None:
        with open(path, "w") as f:
            f.write("period,ndcg,hit_rate,events\n")
            for row in self.result():
                f.write(",".join(str(x) for x in row) + "\n")


class nDCGSink:
    """The reference's evaluation sink with its file formats (T/matrix/factorization/sink/nDCGSink.scala:192-272):
    consumes ``(itemId, timestamp, [(score, itemId)])`` records, scores the rank of the rated item in the
    list (``nDCG = ln 2 / ln(1 + rank)``, 0 if absent), and on ``close()`` writes either the human-readable
    block format or CSV -- ``invokes,avgnDCG,hitrate`` (no periods) / one ``period,invokes,avgnDCG,hitrate``
    line per period of ``periodLength`` timestamp units."""

    def __init__(self, fileName: str, periodLength: int = 86400, append: bool = False, csv: bool = False):
        self.fileName, self.periodLength, self.append, self.csv = fileName, int(periodLength), append, csv
        self.sumnDCG, self.counter, self.hit = 0.0, 0, 0
        self.perDay: Dict[int, List[float]] = {}

    def invoke(self, value) -> None:
        item, timestamp, topk = value[-3], value[-2], value[-1]       # 3-tuples, or 4-tuples led by the user id
        g = 0.0
        for rank, (_score, it) in enumerate(topk, start=1):
            if it == item:
                g = math.log(2.0) / math.log(1.0 + rank)
                break
        self.hit += int(g != 0.0)
        self.sumnDCG += g
        self.counter += 1
        if self.periodLength > 0:
            d = self.perDay.setdefault(int(timestamp // self.periodLength), [0, 0.0, 0])
            d[0] += 1; d[1] += g; d[2] += int(g != 0.0)

    def close(self) -> None:
        avg = self.sumnDCG / self.counter if self.counter else float("nan")
        with open(self.fileName, "a" if self.append else "w") as f:
            if self.csv:
                if self.periodLength <= 0:
                    f.write(f"{self.counter},{avg},{self.hit / self.counter if self.counter else float('nan')}\n")
            else:
                f.write(f"Number of invokes: {self.counter}\nSum nDCG: {self.sumnDCG}\nAvg nDCG: {avg}\n"
                        f"Hit: {self.hit}\n\n")
            if self.periodLength > 0:
                for day in sorted(self.perDay):
                    inv, tot, hit = self.perDay[day]
                    if self.csv:
                        f.write(f"{day},{inv},{tot / inv},{hit / inv}\n")
                    else:
                        f.write(f"Period {day}\n\t:Number of invokes: {inv}\n\t:Sum nDCG: {tot}\n"
                                f"\t:Avg nDCG: {tot / inv}\n\t:Hit: {hit}\n\n")


def _drain(topK, sink: nDCGSink) -> nDCGSink:
    for rec in topK:
        sink.invoke(rec)
    sink.close()
    return sink


def nDCGToFile(topK, fileName: str, periodLength: int = 0) -> nDCGSink:
    """Human-readable totals (+ per-period blocks when ``periodLength > 0``), nDCGSink.scala:42-80."""
    return _drain(topK, nDCGSink(fileName, periodLength, False, False))


def nDCGToCsv(topK, fileName: str) -> nDCGSink:
    """Appends one ``invokes,avgnDCG,hitrate`` line (nDCGSink.scala:91-108)."""
    return _drain(topK, nDCGSink(fileName, 0, True, True))


def nDCGPeriodsToCsv(topK, fileName: str, periodLength: int, append: bool = False) -> nDCGSink:
    """One ``period,invokes,avgnDCG,hitrate`` line per period (nDCGSink.scala:123-147)."""
    return _drain(topK, nDCGSink(fileName, periodLength, append, True))


class Counters:
    """Thread-safe named counters / gauges (pulls, pushes, credit stalls, ring occupancy, ...)."""

    def __init__(self):
        self._v: Dict[str, float] = defaultdict(float)
        self._lock = threading.Lock()

    def inc(self, name: str, by: float = 1) -> None:
        with self._lock:
            self._v[name] += by

    def set(self, name: str, value: float) -> None:
        with self._lock:
            self._v[name] = value

    def snapshot(self) -> Dict[str, float]:
        with self._lock:
            return dict(self._v)


    def prometheus_text(self, prefix: str = "fps_b200_") -> str:
        """Counters in the Prometheus text exposition format (one ``name value`` line each)."""
        lines = []
        for k, v in sorted(self.snapshot().items()):
            name = prefix + "".join(ch if ch.isalnum() else "_" for ch in k)
            lines.append(f"{name} {v:g}")
        return "\n".join(lines) + ("\n" if lines else "")


GLOBAL = Counters()
