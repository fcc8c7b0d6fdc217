#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_workloads.py -x -q -k "item_cache or replica or skipgram" 2>&1 | tail -3
for ec in 0 1 2; do
  FPS_EXCHANGE_CTAS=$ec timeout 200 python bench.py --steps 40 --warmup 8 --item-cache on --sync-every 4 2> gpurun_out/bench30.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=1 cache exchange_ctas=$ec', round(d['value']/1e9,3), round(d['ms_per_step'],4), round(d['e2e']['value']/1e9,3))"
done
