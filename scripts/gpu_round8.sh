#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for f in tests/test_gpu_sketch.py tests/test_gpu_topk.py tests/test_gpu_workloads.py tests/test_gpu_kernels.py tests/test_gpu_pa.py; do
  timeout 240 python -m pytest $f -x -q --timeout 100 --timeout-method=thread > gpurun_out/p8_$(basename $f .py).log 2>&1; echo "$f rc=$?"; tail -4 gpurun_out/p8_$(basename $f .py).log
done
timeout 150 python -m pytest tests/test_gpu_rings.py -x -v -s --timeout 40 --timeout-method=thread > gpurun_out/p8_rings.log 2>&1; echo "rings rc=$?"; tail -40 gpurun_out/p8_rings.log
timeout 200 python benchmarks/topk_bench.py --queries 2048 > gpurun_out/topk_bench.json 2> gpurun_out/topk_bench.err; cat gpurun_out/topk_bench.json; tail -3 gpurun_out/topk_bench.err
