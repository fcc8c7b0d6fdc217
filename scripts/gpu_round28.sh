#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_topk.py -x -q 2>&1 | tail -3
timeout 200 python benchmarks/topk_bench.py --skew 1.0 > gpurun_out/topk_skew.json 2> gpurun_out/topk_skew.err; cat gpurun_out/topk_skew.json; tail -2 gpurun_out/topk_skew.err
timeout 200 python benchmarks/topk_bench.py > gpurun_out/topk_uniform.json 2> gpurun_out/topk_uniform.err; cat gpurun_out/topk_uniform.json; tail -2 gpurun_out/topk_uniform.err
for v in 0 3 5 1; do
  FPS_MF_REG_VARIANT=$v timeout 200 python bench.py --steps 20 --warmup 5 2> gpurun_out/bench28.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('variant=$v', round(d['value']/1e9,3), round(d['ms_per_step'],4), round(d['e2e']['value']/1e9,3))"
done
