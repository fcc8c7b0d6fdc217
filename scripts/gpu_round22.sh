#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_pa.py -x -q 2>&1 | tail -5
timeout 200 python benchmarks/pa_bench.py > gpurun_out/pa_bench.json 2> gpurun_out/pa_bench.err; cat gpurun_out/pa_bench.json; tail -2 gpurun_out/pa_bench.err
