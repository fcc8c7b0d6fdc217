#!/bin/bash
# round 2, call 3 (1 GPU): device backend tests, replica tests, full suite, quality sweep with 8 ranks sharing the GPU
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_device_backend.py -x -q > gpurun_out/c3_backend.log 2>&1; echo "backend rc=$?"; tail -15 gpurun_out/c3_backend.log
timeout 1500 python -m pytest tests/ -x -q -m gpu --deselect tests/test_device_backend.py > gpurun_out/c3_suite.log 2>&1; echo "suite rc=$?"; tail -8 gpurun_out/c3_suite.log
timeout 200 python bench.py --steps 100 --warmup 5 --quality-updates-per-user 0 > gpurun_out/c3_bench_n1.json 2> gpurun_out/c3_bench_n1.err; echo "bench rc=$?"; python -c "import json;d=json.load(open('gpurun_out/c3_bench_n1.json'));print(d['value']/1e9, d['e2e']['value']/1e9)"
FPS_SHARE_GPU=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29731 benchmarks/quality_sweep.py --sync 1,2,4,8 --updates-per-user 400 > gpurun_out/c3_quality_n8.json 2> gpurun_out/c3_quality_n8.err; echo "quality rc=$?"; cat gpurun_out/c3_quality_n8.json; tail -3 gpurun_out/c3_quality_n8.err
