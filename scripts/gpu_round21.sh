#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_topk.py -x -q 2>&1 | tail -5
timeout 200 python benchmarks/topk_bench.py --skew 1.0 > gpurun_out/topk_skew.json 2> gpurun_out/topk_skew.err; cat gpurun_out/topk_skew.json; tail -2 gpurun_out/topk_skew.err
timeout 200 python benchmarks/topk_bench.py > gpurun_out/topk_uniform.json 2> gpurun_out/topk_uniform.err; cat gpurun_out/topk_uniform.json; tail -2 gpurun_out/topk_uniform.err
