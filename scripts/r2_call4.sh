#!/bin/bash
# round 2, call 4 (1 GPU): everything new since call 1, strict timeouts everywhere
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; timeout 420 python -m pytest "$@" -x -q --timeout 150 > gpurun_out/c4_$name.log 2>&1; echo "$name rc=$? $(tail -1 gpurun_out/c4_$name.log)"; }
run backend tests/test_device_backend.py
run output tests/test_gpu_output.py
run sketch tests/test_gpu_sketch.py
run workloads tests/test_gpu_workloads.py
run kernels tests/test_gpu_kernels.py -m gpu
run topk tests/test_gpu_topk.py -m gpu
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q --timeout 430 > gpurun_out/c4_multi.log 2>&1; echo "multi rc=$? $(tail -1 gpurun_out/c4_multi.log)"
run rest tests/test_gpu_pa.py tests/test_gpu_rings.py -m gpu
FPS_SHARE_GPU=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29741 benchmarks/quality_sweep.py --sync 1,2,4 --updates-per-user 400,600,800,1200 --skip-direct > gpurun_out/c4_quality_curves_n8.json 2> gpurun_out/c4_quality_curves_n8.err; echo "quality rc=$?"; cat gpurun_out/c4_quality_curves_n8.json
timeout 200 python bench.py --steps 100 --warmup 5 --quality-updates-per-user 0 > gpurun_out/c4_bench_n1.json 2> gpurun_out/c4_bench_n1.err; echo "bench rc=$?"; python -c "import json;d=json.load(open('gpurun_out/c4_bench_n1.json'));print(d['value']/1e9, d['e2e']['value']/1e9)"
grep -h "Error\|error\|FAILED\|assert" gpurun_out/c4_*.log | head -40
