#!/bin/bash
# what the driver runs at round end, on one GPU
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/ -x -q -m gpu > gpurun_out/final_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -4 gpurun_out/final_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final_smoke.log
timeout 300 python bench.py --impl reference > gpurun_out/final_bench_ref.json 2>&1; cat gpurun_out/final_bench_ref.json
timeout 400 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; cat gpurun_out/final_bench.json; tail -2 gpurun_out/final_bench.err
timeout 200 python benchmarks/workloads_bench.py > gpurun_out/workloads_n1b.json 2> gpurun_out/workloads_n1b.err; grep "^{" gpurun_out/workloads_n1b.json
