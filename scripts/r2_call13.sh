#!/bin/bash
# round 2, call 13 (8 GPUs): message tier with 64 lanes (512 rings per GPU) across NVLink
mkdir -p gpurun_out
timeout 75 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29795 benchmarks/message_tier_bench.py --lanes 64 --iters 2 > gpurun_out/c13_msg_tier_n8_l64.json 2> gpurun_out/c13_msg_tier_n8_l64.err; echo "msg rc=$?"; cat gpurun_out/c13_msg_tier_n8_l64.json | cut -c1-1600; grep -E "Error" gpurun_out/c13_msg_tier_n8_l64.err | head -3
