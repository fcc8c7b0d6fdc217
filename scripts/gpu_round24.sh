#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for hints in 0 1; do for bb in 33554432 16777216 8388608 4194304; do
  FPS_L2_HINTS=$hints FPS_BLOCK_BYTES=$bb timeout 200 python bench.py --steps 20 --warmup 5 --item-blocking on 2> gpurun_out/bench24.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('hints=$hints block=$bb', round(d['value']/1e9,3), round(d['ms_per_step'],4), round(d['e2e']['value']/1e9,3), d['config']['item_blocking'])"
done; done
timeout 120 python -m pytest tests/test_gpu_kernels.py -x -q -k "matches_reference or blocking" 2>&1 | tail -2
