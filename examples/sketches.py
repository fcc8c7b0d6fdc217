"""Streaming sketches: which words co-occur with 'cat'?"""
import sys, os, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fps_b200.models.sketch import jobs as S
from fps_b200.models.sketch.hashing import java_string_hash

r = random.Random(0)
vocab = ["cat", "dog", "fish", "bird", "tree", "car"]
h2w = {java_string_hash(w): w for w in vocab}
tweets = []
for i in range(400):
    ws = ["cat", "dog"] if r.random() < 0.5 else r.sample(vocab, 2)
    tweets.append((str(1000 + i), ws))
bloom = S.bloomFilter(tweets, 4096, 3, 4, 2, 150)
print("bloom   :", [(h2w[k], round(s)) for s, k in dict(S.bloomPredict([(0, "cat")], bloom, 4096, 3, 3, 2, 2, 16, 150))[0]])
tow = S.tugOfWar(tweets, 256, 4, 2, 150)
print("tug-war :", [(h2w[k], round(s)) for s, k in dict(S.tugOfWarPredict([(0, "cat")], tow, 256, 8, 3, 2, 2, 16, 150))[0]])
mh = S.minhash(tweets, 128, 4, 2, 150)
res = dict(S.minhashPredict(["cat"], tweets, mh, 128, 3, 2, 2, 16, 150))
print("minhash :", [(h2w[w], c) for w, c in res[java_string_hash("cat")][:3]])
