"""The reference's 11 sketch experiment mains (M/sketch/**/experiments/*Exp.scala) as one CLI:

    python -m fps_b200.models.sketch.experiments BloomFilterExp <positional args as in the reference>

Positional arguments keep the reference's order.  Train mains read a delimited tweet file
(column 0 = id, column 1 = unix time for the time-aware variants, column 5 = text) and a search-word
file, and write the model as ``id:v1,v2,...`` lines; predict mains read that model back, run the
queries and write ``word - (word,score), ...`` lines.  (The reference's Bloom mains pass
``(numHashes, arraySize)`` where the API expects ``(arraySize, numHashes)``, SURVEY §7.4 -- here the
named values go where they belong.)
"""
from __future__ import annotations

import sys
from typing import Callable, Dict, List

from ...utils import model_io
from . import jobs as S
from .hashing import java_string_hash
from .utils import TimeAwareTweetReader, TweetReader


def _words(path: str) -> List[str]:
    with open(path) as f:
        return [l.strip().lower() for l in f if l.strip()]


def _tweets(path: str, reader) -> List[tuple]:
    out = []
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if line:
                out.extend(reader(line))
    return out


def _write_predictions(path: str, res, hash_to_word: Dict[int, str], key=lambda q: q) -> None:
    with open(path, "w") as f:
        for q, topk in res:
            name = hash_to_word.get(key(q) if not isinstance(q, tuple) else q[0], str(q))
            body = ", ".join(f"({hash_to_word.get(w, w)},{round(s)})" for s, w in topk)
            f.write(f"{name} - {body}\n")


def _queries(path: str):
    return [(java_string_hash(w), w) for w in _words(path)]


# ---- Bloom ------------------------------------------------------------------------------------
def BloomFilterExp(a):
    src, words, delim, model, wP, psP, wait, numHashes, arraySize = a[0], a[1], a[2], a[3], int(a[4]), int(a[5]), float(a[6]), int(a[7]), int(a[8])
    m = S.bloomFilter(_tweets(src, TweetReader(delim, _words(words))), arraySize, numHashes, wP, psP, wait)
    model_io.write_text(model, m)


def BloomFilterPredictExp(a):
    model, inModel, search, pred, wP, psP, wait, pullLimit, numHashes, arraySize, K = a[0], a[1], a[2], a[3], int(a[4]), int(a[5]), float(a[6]), int(a[7]), int(a[8]), int(a[9]), int(a[10])
    h2w = {java_string_hash(w): w for w in _words(inModel)}
    m = list(model_io.read_text(model, int, int, as_set=True))
    res = S.bloomPredict(_queries(search), m, arraySize, numHashes, K, wP, psP, pullLimit, wait)
    _write_predictions(pred, res, h2w)


def TimeAwareBloomFilterExp(a):
    src, words, delim, model, wP, psP, wait, numHashes, arraySize, ts, win = a[0], a[1], a[2], a[3], int(a[4]), int(a[5]), float(a[6]), int(a[7]), int(a[8]), int(a[9]), int(a[10])
    m = S.timeAwareBloomFilter(_tweets(src, TimeAwareTweetReader(delim, _words(words), ts, win)), arraySize, numHashes, wP, psP, wait)
    model_io.write_text(model, m)


def TimeAwareBloomPredictExp(a):
    model, inModel, search, pred, wP, psP, wait, pullLimit, numHashes, arraySize, K = a[0], a[1], a[2], a[3], int(a[4]), int(a[5]), float(a[6]), int(a[7]), int(a[8]), int(a[9]), int(a[10])
    h2w = {java_string_hash(w): w for w in _words(inModel)}
    m = list(model_io.read_text(model, int, int, as_set=True))
    res = S.timeAwareBloomPredict(_queries(search), m, arraySize, numHashes, K, wP, psP, pullLimit, wait)
    _write_predictions(pred, res, h2w)


# ---- tug of war -----------------------------------------------------------------------------------
def TugOfWarExp(a):
    src, words, delim, model, wP, psP, wait, numHashes = a[0], a[1], a[2], a[3], int(a[4]), int(a[5]), float(a[6]), int(a[7])
    model_io.write_text(model, S.tugOfWar(_tweets(src, TweetReader(delim, _words(words))), numHashes, wP, psP, wait))


def TugOfWarPredictExp(a):
    model, inModel, search, pred, wP, psP, wait, pullLimit, numHashes, numMeans, K = a[0], a[1], a[2], a[3], int(a[4]), int(a[5]), float(a[6]), int(a[7]), int(a[8]), int(a[9]), int(a[10])
    h2w = {java_string_hash(w): w for w in _words(inModel)}
    m = [(k, v if isinstance(v, list) else [v]) for k, v in model_io.read_text(model, int, int)]
    res = S.tugOfWarPredict(_queries(search), m, numHashes, numMeans, K, wP, psP, pullLimit, wait)
    _write_predictions(pred, res, h2w)


def TimeAwareToWExp(a):
    src, words, delim, model, wP, psP, wait, numHashes, ts, win = a[0], a[1], a[2], a[3], int(a[4]), int(a[5]), float(a[6]), int(a[7]), int(a[8]), int(a[9])
    m = S.timeAwareTugOfWar(_tweets(src, TimeAwareTweetReader(delim, _words(words), ts, win)), numHashes, wP, psP, wait)
    model_io.write_text(model, m)


def TimeAwareToWPredictExp(a):
    model, inModel, search, pred, wP, psP, wait, pullLimit, numHashes, numMeans, K = a[0], a[1], a[2], a[3], int(a[4]), int(a[5]), float(a[6]), int(a[7]), int(a[8]), int(a[9]), int(a[10])
    h2w = {java_string_hash(w): w for w in _words(inModel)}
    m = [(k, v if isinstance(v, list) else [v]) for k, v in model_io.read_text(model, int, int)]
    res = S.timeAwareTugOfWarPredict(_queries(search), m, numHashes, numMeans, K, wP, psP, pullLimit, wait)
    _write_predictions(pred, res, h2w)


# ---- MinHash ----------------------------------------------------------------------------------
def MinHashExp(a):
    src, words, delim, model, wP, psP, wait, numHashes = a[0], a[1], a[2], a[3], int(a[4]), int(a[5]), float(a[6]), int(a[7])
    model_io.write_text(model, S.minhash(_tweets(src, TweetReader(delim, _words(words))), numHashes, wP, psP, wait))


def MinHashVariationExp(a):
    src, words, delim, model, wP, psP, wait, numHashes = a[0], a[1], a[2], a[3], int(a[4]), int(a[5]), float(a[6]), int(a[7])
    model_io.write_text(model, S.minhashVariation(_tweets(src, TweetReader(delim, _words(words))), numHashes, wP, psP, wait))


def MinHashPredictExp(a):
    model, inModel, search, src, delim, pred, wP, psP, wait, pullLimit, numHashes, K = a[0], a[1], a[2], a[3], a[4], a[5], int(a[6]), int(a[7]), float(a[8]), int(a[9]), int(a[10]), int(a[11])
    allw = _words(inModel)
    h2w = {java_string_hash(w): w for w in allw}
    m = [(k, v if isinstance(v, list) else [v]) for k, v in model_io.read_text(model, int, int)]
    res = S.minhashPredict(_words(search), _tweets(src, TweetReader(delim, allw)), m, numHashes, K, wP, psP, pullLimit, wait)
    with open(pred, "w") as f:
        for q, lst in res:
            f.write(f"{h2w.get(q, q)} - " + ", ".join(f"({h2w.get(w, w)},{c})" for w, c in lst[:K]) + "\n")


MAINS: Dict[str, Callable] = {f.__name__: f for f in [
    BloomFilterExp, BloomFilterPredictExp, TimeAwareBloomFilterExp, TimeAwareBloomPredictExp, TugOfWarExp,
    TugOfWarPredictExp, TimeAwareToWExp, TimeAwareToWPredictExp, MinHashExp, MinHashVariationExp,
    MinHashPredictExp]}


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] not in MAINS:
        print("usage: experiments <" + "|".join(MAINS) + "> args...", file=sys.stderr)
        return 2
    MAINS[argv[0]](argv[1:])
    return 0


if __name__ == "__main__":
    sys.exit(main())
