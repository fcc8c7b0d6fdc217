#!/bin/bash
# round 2, call 10 (8 GPUs): final confirmation -- bench exactly as the driver runs it (defaults), message tier with
# lane-per-message batches across NVLink
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift
  timeout 240 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29792 bench.py --gpus 8 $BARGS > gpurun_out/c10_$name.json 2> gpurun_out/c10_$name.err
  echo "$name rc=$? $(python -c "import json;d=json.load(open('gpurun_out/c10_$name.json'));print(round(d['value']/1e9,2),'G', round(d['ms_per_step'],4),'ms e2e',round(d['e2e']['value']/1e9,2), 'direct', d.get('value_direct') and round(d['value_direct']['value']/1e9,2), 'fp64', d.get('value_fp64') and round(d['value_fp64']['value']/1e9,2), 'q', d['config'].get('quality'))" 2>&1 | tail -1)"; }
BARGS="--steps 400 --warmup 10" run default X=1
timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29791 benchmarks/message_tier_bench.py --lanes 16 > gpurun_out/c10_msg_tier_n8.json 2> gpurun_out/c10_msg_tier_n8.err; echo "msg rc=$?"; cat gpurun_out/c10_msg_tier_n8.json | cut -c1-1600; grep -E "Error" gpurun_out/c10_msg_tier_n8.err | head -3
