"""psOfflineMF on the host tier (exact reference semantics), RMSE on the training ratings."""
import sys, os, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from fps_b200.models.mf.common import Rating
from fps_b200.models.mf.offline import psOfflineMF

r = random.Random(0)
ratings = [Rating(r.randrange(50), r.randrange(40), r.random()) for _ in range(400)]
out = psOfflineMF(ratings, numFactors=16, rangeMin=0.0, rangeMax=0.25, learningRate=0.05, iterations=20,
                  pullLimit=32, workerParallelism=4, psParallelism=4, iterationWaitTime=300,
                  seed=1, plain_residual=True)
users = dict(out.worker_outputs()); items = dict(out.ps_outputs())
rmse = (sum((x.rating - float(np.dot(users[x.user], items[x.item]))) ** 2 for x in ratings) / len(ratings)) ** 0.5
print(f"{len(users)} users, {len(items)} items, train RMSE {rmse:.3f}")
