"""Sketches: helper equations from T/sketch/UtilsTest.scala:11-44 plus end-to-end train/predict
jobs (untested in the reference)."""
import math

from fps_b200.models.sketch import jobs as S
from fps_b200.models.sketch.hashing import hash64, java_string_hash, murmur3_32
from fps_b200.models.sketch.utils import TimeAwareTweetReader, TweetReader, bloomEq, bloomUnion, dotProduct

WAIT = 150
_W = [["cat", "dog"], ["cat", "dog"], ["cat", "dog", "fish"], ["cat", "dog"], ["bird"], ["bird", "fish"],
      ["cat", "dog"], ["fish"], ["cat", "dog"], ["bird"], ["cat", "dog"], ["cat", "fish"]] * 3
TWEETS = [(str(100 + i), ws) for i, ws in enumerate(_W)]


def test_reference_utils_equations():
    assert dotProduct([1, 2, 3], [4, 5, 6]) == 32
    assert abs(bloomEq(20, 3, 4) - 1.487623) < 1e-5
    assert abs(bloomUnion(20, 3, {1, 2, 3, 4}, {3, 4, 5, 6}) - 2.377832) < 1e-5
    assert abs(bloomUnion(20, 3, [1, 2, 3, 4], [3, 4, 5, 6]) - 2.377832) < 1e-5


def test_hashes():
    assert java_string_hash("hello") == 99162322 and java_string_hash("") == 0
    assert java_string_hash("polygenelubricants") == -2147483648   # famous Int.MinValue hash
    assert murmur3_32(b"", 0) == 0 and murmur3_32(b"hello", 0) == 613153351
    assert hash64(1, 0) != hash64(1, 1) and 0 <= hash64(123, 5) < 2 ** 64


def test_tweet_readers():
    line = "42|3600|x|x|x|The CAT sat on a Dog"
    assert TweetReader("|", ["cat", "dog"])(line) == [("42", ["cat", "dog"])]
    assert TimeAwareTweetReader("|", ["cat"], 0, 1)(line) == [("42", ["cat"], 1)]
    assert TweetReader("|", ["zebra"])(line) == []


def test_bloom_train_and_predict_cooccurrence_ranking():
    model = S.bloomFilter(TWEETS, 256, 3, 3, 2, WAIT)
    keys = {k for k, _ in model}
    assert keys == {java_string_hash(w) for w in ["cat", "dog", "fish", "bird"]}
    res = dict(S.bloomPredict([(0, "cat")], model, 256, 3, 2, 2, 2, 10, WAIT))
    top = [k for _, k in res[0]]
    assert set(top) == {java_string_hash("cat"), java_string_hash("dog")}


def test_time_aware_bloom():
    src = [(t, ws, i % 2) for i, (t, ws) in enumerate(TWEETS)]
    model = S.timeAwareBloomFilter(src, 256, 3, 2, 2, WAIT)
    assert all(isinstance(k, tuple) and k[1] in (0, 1) for k, _ in model)
    res = S.timeAwareBloomPredict([(7, "cat")], model, 256, 3, 2, 2, 2, 10, WAIT)
    assert sorted(q for q, _ in res) == [(7, 0), (7, 1)]


def test_tug_of_war_train_predict():
    model = S.tugOfWar(TWEETS, 128, 3, 2, WAIT)
    cat = dict(model)[java_string_hash("cat")]
    assert len(cat) == 128 and all(abs(c) <= 30 for c in cat)
    res = dict(S.tugOfWarPredict([(1, "cat")], model, 128, 4, 2, 2, 2, 10, WAIT))
    assert {k for _, k in res[1]} == {java_string_hash("cat"), java_string_hash("dog")}
    ta = S.timeAwareTugOfWar([(t, ws, 0) for t, ws in TWEETS], 64, 2, 2, WAIT)
    r2 = S.timeAwareTugOfWarPredict([(1, "dog")], ta, 64, 4, 2, 2, 2, 10, WAIT)
    assert r2 and r2[0][0] == (1, 0)
    assert abs(S.median_of_means([1, -1, 1, 1], [1, -1, 1, 1], 4, 2) - 1.0) < 1e-12


def test_minhash_train_predict_and_variation():
    model = S.minhash(TWEETS, 64, 3, 2, WAIT)
    sig = dict(model)
    cat, dog, bird = (sig[java_string_hash(w)] for w in ["cat", "dog", "bird"])
    jac = lambda a, b: sum(x == y for x, y in zip(a, b)) / 64
    assert jac(cat, dog) > 0.5 > jac(cat, bird)
    res = dict(S.minhashPredict(["cat"], TWEETS, model, 64, 3, 2, 2, 10, WAIT))
    ranked = [w for w, _ in res[java_string_hash("cat")]]
    assert ranked[0] in (java_string_hash("cat"), java_string_hash("dog"))
    var = dict(S.minhashVariation(TWEETS, 16, 2, 2, WAIT))
    tweet_ids = {int(t) for t, _ in TWEETS}
    assert all(t in tweet_ids for t in var[java_string_hash("cat")])
    quirk = dict(S.minhashVariation(TWEETS[:1], 4, 1, 1, WAIT, seed_with_word_id=True))
    assert quirk[java_string_hash("cat")] == [java_string_hash("cat")] * 4
