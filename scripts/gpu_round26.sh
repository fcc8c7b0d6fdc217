#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_workloads.py -x -q -k "item_cache or replica or skipgram" 2>&1 | tail -3
for v in fused 2phase; do for se in 1 4; do
  FPS_CACHE_SYNC=$v timeout 200 python bench.py --steps 20 --warmup 5 --item-cache on --sync-every $se 2> gpurun_out/bench26.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v sync_every=$se', round(d['value']/1e9,3), round(d['ms_per_step'],4), round(d['e2e']['value']/1e9,3), d['gpu_launches'])"
done; done
