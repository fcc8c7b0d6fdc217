"""Offline evaluation tools of the reference's L7 layer: what its notebooks and its
``scripts/evaluation.sh`` do around the jobs, as importable functions and a small CLI.

* ``split_sessions`` / ``session_stats`` -- ``Notebooks/Data_Manipulation.ipynb``: cut a ``<time> <user> <item>``
  event log into ``train_days`` of training and the following ``test_days`` of test events, with the counts
  the notebook prints (events, users, items, users in both parts, quantiles of test events per user).
* ``read_vector_map`` + ``precision_recall_at_k`` -- ``Notebooks/Tester.ipynb``: load the ``id;value`` vector
  maps written by the online-MF main (``models/mf/experiments.py``), recommend the top-``k`` unseen items to
  every active user by dot product and count hits against the test events.
* ``rank_correlations`` / ``evaluate_predictions`` -- ``scripts/evaluation.sh``: compare the sketch jobs'
  predicted co-occurrence lists with exact counts; Pearson, Spearman, Kendall tau and weighted Kendall tau
  per query word, averaged over the words with a defined value (the script drops NaN rows).

    python -m fps_b200.utils.evaluation split <events> <train.out> <test.out> [train_days] [test_days]
    python -m fps_b200.utils.evaluation recall <userVectors> <itemVectors> <train> <test> [k]
    python -m fps_b200.utils.evaluation correlate <exact predictions file> <sketch predictions file>
"""
from __future__ import annotations

import math
import re
import sys
from collections import Counter, defaultdict
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

DAY = 86400


# ---- Data_Manipulation.ipynb -------------------------------------------------------------------------------
def read_events(path: str) -> List[Tuple[int, int, int]]:
    """``<time> <user> <item>`` per line (blank separated)."""
    out = []
    with open(path) as f:
        for line in f:
            p = line.split()
            if len(p) >= 3:
                out.append((int(p[0]), int(p[1]), int(p[2])))
    return out


def split_sessions(events: Sequence[Tuple[int, int, int]], train_days: int = 31, test_days: int = 14):
    """Train = events up to ``t0 + train_days`` days (``t0`` = time of the first event), test = the
    ``test_days`` days after that; later events are dropped.  Returns ``(train, test)``; test events are
    ``(user, item)`` pairs sorted by user, as the notebook writes them."""
    if not events:
        return [], []
    t0 = events[0][0]
    t1 = t0 + DAY * int(train_days)
    t2 = t1 + DAY * int(test_days)
    train = [e for e in events if e[0] <= t1]
    test = sorted(((u, i) for (t, u, i) in events if t1 < t <= t2), key=lambda x: x[0])
    return train, test


def session_stats(train, test) -> Dict[str, float]:
    tu, ti = {u for _, u, _ in train}, {i for _, _, i in train}
    per_user = Counter(u for u, _ in test)
    counts = np.array(sorted(per_user.values()), dtype=np.float64) if per_user else np.zeros(1)
    return {"train_events": len(train), "train_users": len(tu), "train_items": len(ti),
            "max_item_id": max(ti) if ti else 0, "test_events": len(test), "test_users": len(per_user),
            "users_in_both": len(tu & set(per_user)), "test_events_per_user_min": float(counts.min()),
            "test_events_per_user_max": float(counts.max()), "test_events_per_user_mean": float(counts.mean()),
            **{f"test_events_per_user_q{int(q * 100)}": float(np.quantile(counts, q)) for q in (0.5, 0.75, 0.85)}}


def write_split(train, test, train_path: str, test_path: str) -> None:
    with open(train_path, "w") as f:
        for t, u, i in train:
            f.write(f"{t} {u} {i}\n")
    with open(test_path, "w") as f:
        for u, i in test:
            f.write(f"{u} {i}\n")


# ---- Tester.ipynb ---------------------------------------------------------------------------------------------
def read_vector_map(path: str) -> Dict[int, np.ndarray]:
    """``id;value`` lines, components in order."""
    acc: Dict[int, List[float]] = defaultdict(list)
    with open(path) as f:
        for line in f:
            line = line.strip()
            if line:
                k, v = line.split(";", 1)
                acc[int(k)].append(float(v))
    return {k: np.asarray(v, dtype=np.float64) for k, v in acc.items()}


def precision_recall_at_k(user_vectors: Dict[int, np.ndarray], item_vectors: Dict[int, np.ndarray],
                          train: Iterable[Tuple[int, int, int]], test: Iterable[Tuple[int, int]], k: int = 5,
                          max_item_frequency: Optional[int] = None) -> Dict[str, float]:
    """Top-``k`` unseen items by ``u . v`` for every active user (seen in train and test, has a vector), scored
    against that user's test items.  ``max_item_frequency``: keep only train events of items seen at most that
    many times, as the notebook's filtering cell does (it keeps items with frequency <= 5); ``None`` = no
    filter.  Items are candidates when they occur in both parts and have a vector."""
    train = list(train)
    if max_item_frequency is not None:
        freq = Counter(i for _, _, i in train)
        train = [e for e in train if freq[e[2]] <= max_item_frequency]
    seen: Dict[int, set] = defaultdict(set)
    for _, u, i in train:
        seen[u].add(i)
    wanted: Dict[int, List[int]] = defaultdict(list)
    for u, i in test:
        wanted[u].append(i)
    train_items = {i for _, _, i in train}
    items = sorted(i for i in {i for _, i in test} & train_items if i in item_vectors)
    users = sorted(u for u in seen if u in wanted and u in user_vectors)
    if not users or not items:
        return {"users": len(users), "items": len(items), "hits": 0, "precision": 0.0, "recall": 0.0}
    V = np.stack([item_vectors[i] for i in items])
    col = {i: c for c, i in enumerate(items)}
    hits, n_test = 0, 0
    for u in users:
        s = V @ user_vectors[u]
        for i in seen[u]:
            c = col.get(i)
            if c is not None:
                s[c] = -np.inf
        kk = min(k, len(items))
        top = np.argpartition(-s, kk - 1)[:kk]
        rec = {items[c] for c in top if np.isfinite(s[c])}
        hits += len(rec & set(wanted[u]))
        n_test += len(wanted[u])
    return {"users": len(users), "items": len(items), "hits": hits, "precision": hits / (len(users) * float(k)),
            "recall": hits / float(n_test)}


# ---- scripts/evaluation.sh ---------------------------------------------------------------------------------------
_PAIR = re.compile(r"\(([^,()]+),\s*(-?[0-9.eE+]+)\)")


def read_predictions(path: str) -> Dict[str, Dict[str, float]]:
    """``word - (other,score), (other,score), ...`` lines (what the predict mains write)."""
    out: Dict[str, Dict[str, float]] = {}
    with open(path) as f:
        for line in f:
            if " - " not in line:
                continue
            q, body = line.split(" - ", 1)
            out[q.strip()] = {w.strip(): float(s) for w, s in _PAIR.findall(body)}
    return out


def rank_correlations(exact: Dict[str, float], predicted: Dict[str, float]) -> Dict[str, float]:
    """Correlation of two score maps over the union of their keys (missing = 0): Pearson, Spearman, Kendall tau,
    weighted Kendall tau.  NaN when a side is constant."""
    from scipy import stats

    keys = sorted(set(exact) | set(predicted))
    a = np.array([exact.get(k, 0.0) for k in keys], dtype=np.float64)
    b = np.array([predicted.get(k, 0.0) for k in keys], dtype=np.float64)
    if len(keys) < 2 or a.std() == 0.0 or b.std() == 0.0:
        nan = float("nan")
        return {"pearson": nan, "spearman": nan, "kendall": nan, "weighted_kendall": nan, "n": len(keys)}
    return {"pearson": float(stats.pearsonr(a, b)[0]), "spearman": float(stats.spearmanr(a, b)[0]),
            "kendall": float(stats.kendalltau(a, b)[0]), "weighted_kendall": float(stats.weightedtau(a, b)[0]),
            "n": len(keys)}


def evaluate_predictions(exact: Dict[str, Dict[str, float]], predicted: Dict[str, Dict[str, float]],
                         words: Optional[Iterable[str]] = None) -> Dict[str, object]:
    """Per-word correlations + their averages over the words where every coefficient is defined
    (``words``: restrict to a word list, like the script's three frequency classes)."""
    names = ("pearson", "spearman", "kendall", "weighted_kendall")
    per_word = {}
    for w in (list(words) if words is not None else sorted(exact)):
        if w in exact and w in predicted:
            per_word[w] = rank_correlations(exact[w], predicted[w])
    ok = [c for c in per_word.values() if not any(math.isnan(c[n]) for n in names)]
    avg = {n: (sum(c[n] for c in ok) / len(ok) if ok else float("nan")) for n in names}
    return {"per_word": per_word, "average": avg, "words": len(per_word), "words_defined": len(ok)}


def main(argv: List[str]) -> int:
    if len(argv) >= 4 and argv[0] == "split":
        ev = read_events(argv[1])
        tr, te = split_sessions(ev, int(argv[4]) if len(argv) > 4 else 31, int(argv[5]) if len(argv) > 5 else 14)
        write_split(tr, te, argv[2], argv[3])
        print(session_stats(tr, te))
        return 0
    if len(argv) >= 5 and argv[0] == "recall":
        test = []
        with open(argv[4]) as f:
            for line in f:
                p = line.split()
                if len(p) >= 2:
                    test.append((int(p[0]), int(p[1])))
        print(precision_recall_at_k(read_vector_map(argv[1]), read_vector_map(argv[2]), read_events(argv[3]), test,
                                    int(argv[5]) if len(argv) > 5 else 5))
        return 0
    if len(argv) >= 3 and argv[0] == "correlate":
        r = evaluate_predictions(read_predictions(argv[1]), read_predictions(argv[2]))
        print({"average": r["average"], "words": r["words"], "words_defined": r["words_defined"]})
        return 0
    print(__doc__)
    return 2


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
