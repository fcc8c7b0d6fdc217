"""Binary passive-aggressive classifier: train, dump the model, load it back and predict."""
import sys, os, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fps_b200.api import Left, Right
from fps_b200.models.pa.algorithms import PassiveAggressiveBinaryAlgorithm
from fps_b200.models.pa.ps import transformBinary
from fps_b200.models.pa.sparse import SparseVector

backend = "device" if "--device" in sys.argv else "local"
r = random.Random(1)
feats = 400
w_true = [r.gauss(0, 1) for _ in range(feats)]
def example():
    idx = r.sample(range(feats), 60)
    v = SparseVector(idx, [r.gauss(0, 1) for _ in idx], feats)
    return v, sum(w_true[i] * x for i, x in v.activeIterator()) > 0
data = [example() for _ in range(1050)]
algo = PassiveAggressiveBinaryAlgorithm.buildPAI(1.0)
trained = transformBinary()([Left(d) for d in data[:1000]] * 3, 3, 3, algo, 1000, feats, True, 200, backend=backend)
model = list(trained.ps_outputs())
pred = transformBinary(model)([Right((i, v)) for i, (v, _) in enumerate(data[1000:])], 3, 3, algo, 1000, feats, True, 200,
                              backend=backend)
got = {v: bool(p) for v, p in pred.worker_outputs()}
acc = sum(got[v] == y for v, y in data[1000:]) / 50
print(f"backend={backend}: {len(model)} weights, held-out accuracy {acc:.2f}")
