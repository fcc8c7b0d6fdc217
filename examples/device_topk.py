"""Top-K serving on B200: user vectors on the parameter server, tcgen05 scoring against local items."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fps_b200.models.mf.device_topk import DeviceTopK
from fps_b200.store.sharded_table import ShardedTable

torch.cuda.set_device(0)
users = ShardedTable(100_000, 64, seed=1, init_range=(-1, 1))          # parameter server
items = ShardedTable(200_000, 64, seed=2, init_range=(-1, 1))          # this worker's items
queries = torch.randint(0, 100_000, (1024,), device="cuda")
scores, rows = DeviceTopK(items.local).topk(10, q_ids=queries, q_table=users, rescore=True)
print("user", int(queries[0]), "->", rows[0].tolist(), [round(s, 2) for s in scores[0].tolist()])
