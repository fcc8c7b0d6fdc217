#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "bucket or blocking or fit_stream" 2>&1 | tail -3
for i in 1 2; do
timeout 200 python bench.py --steps 50 --warmup 5 2> gpurun_out/bench29.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', round(d['value']/1e9,3), round(d['ms_per_step'],4), round(d['e2e']['value']/1e9,3), d['clocks']['reasons'])"
done
