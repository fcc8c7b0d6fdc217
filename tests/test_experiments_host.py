"""L7 of the reference: the two MF experiment mains, the nDCG sink's file formats, and the offline
evaluation the notebooks / ``scripts/evaluation.sh`` perform."""
import random

import numpy as np
import pytest

from fps_b200.models.mf import experiments as E
from fps_b200.utils import evaluation as V
from fps_b200.utils.metrics import nDCGPeriodsToCsv, nDCGToCsv, nDCGToFile

FAST = dict(workerParallelism=2, psParallelism=2, iterationWaitTime=100)


def _events(n=240, users=12, items=20, seed=3, t0=1_000_000, span_days=3):
    r = random.Random(seed)
    ev = sorted((t0 + r.randrange(span_days * 86400), r.randrange(users), r.randrange(items)) for _ in range(n))
    ev[0] = (t0, ev[0][1], ev[0][2])
    return ev


@pytest.mark.parametrize("backend", ["local", "native"])
def test_online_mf_implicit_main_writes_vector_maps(tmp_path, backend):
    ev = _events()
    src = tmp_path / "data_session_train"
    src.write_text("".join(f"{t} {u} {i}\n" for t, u, i in ev))
    uo, io = tmp_path / "UserVector.map", tmp_path / "ItemVector.map"
    E.OnlineMFImplicit([str(src), str(uo), str(io), backend], **FAST)
    users, items = V.read_vector_map(str(uo)), V.read_vector_map(str(io))
    assert set(users) == {u for _, u, _ in ev}
    assert {i for _, _, i in ev} <= set(items)                 # negatives may add more items
    assert all(v.shape == (10,) and np.isfinite(v).all() for v in list(users.values()) + list(items.values()))
    # the notebook's offline evaluation runs on those files
    tr, te = V.split_sessions(ev, train_days=2, test_days=1)
    res = V.precision_recall_at_k(users, items, tr, te, k=5)
    assert res["users"] > 0 and 0.0 <= res["precision"] <= 1.0 and 0.0 <= res["recall"] <= 1.0


def test_online_learner_and_topk_main_writes_per_day_ndcg(tmp_path):
    ev = _events(n=150, users=8, items=12, span_days=2)
    src = tmp_path / "week_all"
    src.write_text("".join(f"{t},{u},{i}\n" for t, u, i in ev))
    out = tmp_path / "onlineMF_nDCG.csv"
    sink = E.OnlineLearnerAndTopK([str(src), str(out)], K=5, workerK=5, bucketSize=4, **FAST)
    rows = [l.split(",") for l in out.read_text().strip().splitlines()]
    assert sum(int(r[1]) for r in rows) == len(ev) == sink.counter
    assert [int(r[0]) for r in rows] == sorted({t // 86400 for t, _, _ in ev})
    assert all(0.0 <= float(r[2]) <= 1.0 and 0.0 <= float(r[3]) <= 1.0 for r in rows)
    assert E.main(["nope"]) == 2


def test_ndcg_sink_formats(tmp_path):
    recs = [(1, 5, 10, [(0.9, 5), (0.8, 6)]), (2, 6, 20, [(0.9, 5), (0.8, 6)]), (3, 7, 86401, [(0.9, 5)])]
    p = tmp_path / "a.csv"
    nDCGPeriodsToCsv(recs, str(p), 86400)
    l0, l1 = p.read_text().splitlines()
    assert l0.startswith("0,2,") and abs(float(l0.split(",")[2]) - (1 + np.log(2) / np.log(3)) / 2) < 1e-12
    assert l1 == "1,1,0.0,0.0"
    t = tmp_path / "a.txt"
    nDCGToFile([r[1:] for r in recs], str(t), 0)                           # 3-tuples work too; no periods
    assert t.read_text().startswith("Number of invokes: 3\nSum nDCG: ") and "Period" not in t.read_text()
    c = tmp_path / "b.csv"
    nDCGToCsv(recs, str(c)); nDCGToCsv(recs, str(c))                      # appends
    assert len(c.read_text().splitlines()) == 2 and c.read_text().startswith("3,")


def test_session_split_stats_and_recall():
    ev = [(0, 1, 10), (10, 1, 11), (86400 * 31, 2, 10), (86400 * 31 + 1, 1, 12), (86400 * 40, 2, 11),
          (86400 * 46, 2, 13)]
    tr, te = V.split_sessions(ev)
    assert [e[0] for e in tr] == [0, 10, 86400 * 31] and te == [(1, 12), (2, 11)]       # day 46 is dropped
    st = V.session_stats(tr, te)
    assert st["train_events"] == 3 and st["train_users"] == 2 and st["users_in_both"] == 2 and st["test_users"] == 2
    # planted model: user 1 loves item 12 (unseen), user 2 loves item 11 (unseen)
    items = {10: np.array([1.0, 0.0]), 11: np.array([0.0, 1.0]), 12: np.array([0.7, -0.7]), 13: np.array([-1.0, -1.0])}
    users = {1: np.array([1.0, -1.0]), 2: np.array([-0.2, 1.0])}
    train = tr + [(5, 9, 12), (6, 9, 11)]                                 # make 11 and 12 train items as well
    res = V.precision_recall_at_k(users, items, train, te, k=1)
    assert res == {"users": 2, "items": 2, "hits": 2, "precision": 1.0, "recall": 1.0}
    # the notebook's frequency filter keeps train events of RARE items only: here just (user 9, item 12)
    res = V.precision_recall_at_k(users, items, train, te, k=1, max_item_frequency=1)
    assert res["users"] == 0 and res["hits"] == 0
    res = V.precision_recall_at_k(users, items, train, te, k=1, max_item_frequency=2)
    assert res["users"] == 2 and res["hits"] == 2


def test_rank_correlations_and_prediction_files(tmp_path):
    exact = {"a": 10.0, "b": 7.0, "c": 3.0, "d": 1.0}
    same = V.rank_correlations(exact, {k: 2 * v + 1 for k, v in exact.items()})
    assert all(abs(same[n] - 1.0) < 1e-9 for n in ("pearson", "spearman", "kendall", "weighted_kendall"))
    rev = V.rank_correlations(exact, {"a": 1.0, "b": 3.0, "c": 7.0, "d": 10.0})
    assert rev["spearman"] < -0.99 and rev["kendall"] < -0.99
    assert not np.isnan(V.rank_correlations(exact, {"a": 1.0})["pearson"])           # missing keys count as 0
    assert np.isnan(V.rank_correlations({"a": 1.0, "b": 1.0}, {"a": 2.0, "b": 3.0})["pearson"])   # constant side
    ex, pr = tmp_path / "exact.txt", tmp_path / "pred.txt"
    ex.write_text("cat - (dog,10), (fish,4), (bird,1)\nsun - (moon,5), (star,5)\n")
    pr.write_text("cat - (dog,9), (fish,5), (bird,0)\nsun - (moon,2), (star,1)\n")
    r = V.evaluate_predictions(V.read_predictions(str(ex)), V.read_predictions(str(pr)))
    assert r["words"] == 2 and r["words_defined"] == 1 and r["average"]["spearman"] > 0.99   # "sun" is constant: dropped
    assert V.main(["correlate", str(ex), str(pr)]) == 0 and V.main([]) == 2
