#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_topk.py -x -q --timeout 80 --timeout-method=thread > gpurun_out/p14_a.log 2>&1; echo "kernels+topk rc=$?"; grep -E "passed|failed|Error|Mismatch|Greatest|assert" gpurun_out/p14_a.log | head -12
timeout 120 python benchmarks/topk_bench.py --queries 2048 > gpurun_out/topk_bench3.json 2> gpurun_out/topk_bench3.err; cat gpurun_out/topk_bench3.json; tail -2 gpurun_out/topk_bench3.err
timeout 200 python bench.py --steps 200 --warmup 10 --item-cache on --sync-every 2 > gpurun_out/b14_n1_cache.json 2> gpurun_out/b14_n1_cache.err; python -c "import json;d=json.loads(open('gpurun_out/b14_n1_cache.json').read().strip().splitlines()[-1]);print('n1 cache', round(d['value']/1e9,3), d['ms_per_step'], d['e2e']['last_step_mse'])"; tail -2 gpurun_out/b14_n1_cache.err
