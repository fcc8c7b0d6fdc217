#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 120 python -m pytest tests/ -x -q -m gpu > gpurun_out/final_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -3 gpurun_out/final_pytest_gpu.log
timeout 60 python benchmarks/graph_step_bench.py > gpurun_out/graph_step_bench.json 2> gpurun_out/graph_step_bench.err; cat gpurun_out/graph_step_bench.json; tail -2 gpurun_out/graph_step_bench.err
