#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( time CUDA_MODULE_LOADING=EAGER python -c "import torch; x=torch.zeros(4,device='cuda'); torch.cuda.synchronize(); print('eager ok')" ) 2>&1 | grep -E "real|ok"
( time CUDA_MODULE_LOADING=LAZY python -c "import torch; x=torch.zeros(4,device='cuda'); torch.cuda.synchronize(); print('lazy ok')" ) 2>&1 | grep -E "real|ok"
timeout 120 python benchmarks/topk_bench.py --queries 2048 > gpurun_out/topk_bench4.json 2> gpurun_out/topk_bench4.err; cat gpurun_out/topk_bench4.json; tail -2 gpurun_out/topk_bench4.err
timeout 120 python -m pytest tests/test_gpu_topk.py -x -q --timeout 60 --timeout-method=thread 2>&1 | tail -2
timeout 500 python -m pytest tests/test_gpu_pa.py tests/test_gpu_rings.py tests/test_gpu_workloads.py tests/test_gpu_sketch.py -x -q --durations=12 --timeout 120 --timeout-method=thread > gpurun_out/p16_dur.log 2>&1; echo "rc=$?"; grep -E "s call|s setup|passed|failed" gpurun_out/p16_dur.log | head -20
