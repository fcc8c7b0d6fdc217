#!/bin/bash
# round 2, call 6 (2 GPUs): fixes since call 5 -- learner check, cross-GPU Lock-A transactions, message-tier bench, fp64 kernel test
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 python -m pytest tests/test_gpu_kernels.py -x -q --timeout 120 -k "fp64" > gpurun_out/c6_fp64.log 2>&1; echo "fp64 rc=$? $(tail -1 gpurun_out/c6_fp64.log)"
timeout 420 python -m pytest tests/test_gpu_multi.py -x -q --timeout 200 -k "fabric or learner" > gpurun_out/c6_multi.log 2>&1; echo "multi rc=$? $(tail -1 gpurun_out/c6_multi.log)"
grep -E "Error|assert" gpurun_out/c6_multi.log | head -10
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29762 benchmarks/message_tier_bench.py > gpurun_out/c6_msg_tier_n2.json 2> gpurun_out/c6_msg_tier_n2.err; echo "msg rc=$?"; cat gpurun_out/c6_msg_tier_n2.json; grep -E "Error|error" gpurun_out/c6_msg_tier_n2.err | head -5
timeout 120 python benchmarks/message_tier_bench.py > gpurun_out/c6_msg_tier_n1.json 2> gpurun_out/c6_msg_tier_n1.err; echo "msg1 rc=$?"; cat gpurun_out/c6_msg_tier_n1.json
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29763 bench.py --gpus 2 --steps 100 --warmup 5 --quality-updates-per-user 0 > gpurun_out/c6_bench_n2.json 2> gpurun_out/c6_bench_n2.err; echo "bench rc=$?"; python -c "import json;d=json.load(open('gpurun_out/c6_bench_n2.json'));print(d['value']/1e9, d['ms_per_step'], 'fp64', d['value_fp64'], 'direct', d['value_direct'])"
