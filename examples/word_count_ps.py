"""Word count on the parameter server: every word is a parameter id, workers push +1."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from collections import Counter

from fps_b200 import WorkerLogic, addPullLimiter, transform

TEXT = "the quick brown fox jumps over the lazy dog and the fox runs away".split() * 100


class CountWorker(WorkerLogic):
    def onRecv(self, word, ps):
        ps.pull(word)                     # ask for the current count ...

    def onPullRecv(self, word, count, ps):
        ps.push(word, 1)                  # ... and add one (paramUpdate = +)


out = transform(TEXT, addPullLimiter(CountWorker(), 64),
                lambda word: 0, lambda count, delta: count + delta,
                4, 3, 100)                # workerParallelism, psParallelism, iterationWaitTime (ms)
final = {}
for word, count in out.ps_outputs():      # SimplePSLogic emits (id, value) on every push
    final[word] = max(final.get(word, 0), count)
assert final == dict(Counter(TEXT))
print(sorted(final.items(), key=lambda kv: -kv[1])[:5])
