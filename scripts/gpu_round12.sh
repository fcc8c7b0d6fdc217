#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 150 python -m pytest tests/test_gpu_kernels.py -x -q -k "item_cache or wide_rows" --timeout 60 --timeout-method=thread > gpurun_out/p12_a.log 2>&1; echo "cache1gpu rc=$?"; grep -E "passed|failed|Error|Mismatch|Greatest|assert" gpurun_out/p12_a.log | head -12
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29617 tests/mp_device_check.py > gpurun_out/p12_mp.log 2>&1; echo "mp rc=$?"; grep -E "OK|rank0.*(Error|assert|Mismatch|Greatest|!=)" gpurun_out/p12_mp.log | head -8
