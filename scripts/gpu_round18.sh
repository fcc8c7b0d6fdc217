#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/p18_multi.log 2>&1; echo "multi rc=$?"; tail -5 gpurun_out/p18_multi.log | grep -v "OMP\|\*\*\*"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 benchmarks/workloads_bench.py --slots 200000000 > gpurun_out/workloads_n2.json 2> gpurun_out/workloads_n2.err; grep "^{" gpurun_out/workloads_n2.json; grep -v "OMP\|\*\*\*" gpurun_out/workloads_n2.err | tail -3
