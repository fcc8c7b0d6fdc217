#!/bin/bash
# round 2, call 1 (1 GPU): new replica exchange + shared-GPU multi-rank harness, then the full suite, then N=1 bench
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "replica or bucket" > gpurun_out/c1_replica.log 2>&1; echo "replica rc=$?" 
tail -5 gpurun_out/c1_replica.log
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/c1_multi.log 2>&1; echo "multi rc=$?"
tail -15 gpurun_out/c1_multi.log
timeout 1200 python -m pytest tests/ -x -q -m gpu --deselect tests/test_gpu_multi.py > gpurun_out/c1_suite.log 2>&1; echo "suite rc=$?"
tail -5 gpurun_out/c1_suite.log
timeout 300 python bench.py --steps 200 --warmup 10 > gpurun_out/c1_bench_n1.json 2> gpurun_out/c1_bench_n1.err; echo "bench rc=$?"
cat gpurun_out/c1_bench_n1.json
# functional (not perf): two ranks sharing the GPU through the new exchange path
FPS_SHARE_GPU=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 20 --warmup 3 --users 2000000 --items 1000000 --quality-updates-per-user 20 > gpurun_out/c1_bench_shared2.json 2> gpurun_out/c1_bench_shared2.err; echo "shared bench rc=$?"
cat gpurun_out/c1_bench_shared2.json; tail -5 gpurun_out/c1_bench_shared2.err
