"""Sketch update kernels (red.or / red.add.s32 / red.min.u64) vs the host oracle with the same hashes."""
import pytest
import torch

pytestmark = pytest.mark.gpu

_W = [["cat", "dog"], ["cat", "dog"], ["cat", "dog", "fish"], ["cat", "dog"], ["bird"], ["bird", "fish"],
      ["cat", "dog"], ["fish"], ["cat", "dog"], ["bird"], ["cat", "dog"], ["cat", "fish"]] * 5
TWEETS = [(str(1000 + i), ws) for i, ws in enumerate(_W)]


def _by_word(model):
    from fps_b200.models.sketch.hashing import java_string_hash
    d = dict(model)
    return {w: d[java_string_hash(w)] for w in ["cat", "dog", "fish", "bird"]}


def test_bloom_device_equals_oracle_and_ranks_cooccurrence():
    from fps_b200.models.sketch.device import DeviceSketch, bloom_positions64
    from fps_b200.models.sketch.hashing import java_string_hash

    torch.cuda.set_device(0)
    sk = DeviceSketch("bloom", 64, 3, 1024)
    sk.update(TWEETS[:30]); sk.update(TWEETS[30:])
    got = _by_word(sk.model())
    for w in got:
        ref = set()
        for tid, ws in TWEETS:
            if w in ws:
                ref.update(bloom_positions64(int(tid), 3, 1024))
        assert got[w] == frozenset(ref)
    top = sk.query_local("cat", 2)
    assert {k for _, k in top} == {java_string_hash("cat"), java_string_hash("dog")}
    sk.close()


def test_tow_and_minhash_device_equal_oracle():
    from fps_b200.models.sketch.device import DeviceSketch, minhash_packed64, tow_bits64
    from fps_b200.models.sketch.hashing import java_string_hash

    torch.cuda.set_device(0)
    tow = DeviceSketch("tow", 64, 96)
    tow.update(TWEETS)
    got = _by_word(tow.model())
    for w in got:
        ref = [0] * 96
        for tid, ws in TWEETS:
            if w in ws:
                ref = [a + b for a, b in zip(ref, tow_bits64(int(tid), 96))]
        assert got[w] == ref
    assert {k for _, k in tow.query_local("dog", 2)} == {java_string_hash("cat"), java_string_hash("dog")}
    tow.close()
    mh = DeviceSketch("minhash", 64, 32)
    mh.update(TWEETS)
    got = _by_word(mh.model())
    for w in got:
        ref = None
        for tid, ws in TWEETS:
            if w in ws:
                p = minhash_packed64(int(tid), 32)
                ref = p if ref is None else [min(a, b) for a, b in zip(ref, p)]
        assert got[w] == [r & 0xFFFFFFFF for r in ref]
    top = mh.query_local("cat", 2)
    assert top[0][1] == java_string_hash("cat") and top[1][1] == java_string_hash("dog")
    mh.close()
