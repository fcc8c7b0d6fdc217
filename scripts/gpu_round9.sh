#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 150 python -m pytest tests/test_gpu_rings.py -x -v --timeout 40 --timeout-method=thread > gpurun_out/p9_rings.log 2>&1; echo "rings rc=$?"; grep -E "PASS|FAIL|Timeout|Error" gpurun_out/p9_rings.log | head -20
timeout 200 python -m pytest tests/test_gpu_workloads.py tests/test_gpu_kernels.py -x -q --timeout 100 --timeout-method=thread > gpurun_out/p9_rest.log 2>&1; echo "rest rc=$?"; tail -4 gpurun_out/p9_rest.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fps_topk_mma -s 2 -c 1 -o gpurun_out/prof_topk -f python benchmarks/topk_bench.py --queries 512 --items 200000 > gpurun_out/ncu_topk.log 2>&1; echo "ncu rc=$?"
