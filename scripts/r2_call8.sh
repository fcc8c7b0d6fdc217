#!/bin/bash
# round 2, call 8 (8 GPUs): the scaling run -- N=8 bench (default config, as the driver runs it), exchange variants,
# message tier and CTR at 1B slots across NVLink
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { # name, env..., -- args
  name=$1; shift
  timeout 300 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29781 bench.py --gpus 8 --steps 200 --warmup 10 $BARGS > gpurun_out/c8_$name.json 2> gpurun_out/c8_$name.err
  echo "$name rc=$? $(python -c "import json;d=json.load(open('gpurun_out/c8_$name.json'));print(round(d['value']/1e9,2),'G', round(d['ms_per_step'],4),'ms e2e',round(d['e2e']['value']/1e9,2), 'direct', d.get('value_direct') and round(d['value_direct']['value']/1e9,2), 'fp64', d.get('value_fp64') and round(d['value_fp64']['value']/1e9,2), 'q', d['config'].get('quality'), 'exch', d['config'].get('exchange'))" 2>&1 | tail -1)"
}
BARGS="" run default FPS_EXCHANGE_TIMING=1
BARGS="--no-direct --no-fp64 --quality-updates-per-user 0" run ctas64 FPS_EXCHANGE_CTAS=64
BARGS="--no-direct --no-fp64 --quality-updates-per-user 0 --sync-every 2" run se2 X=1
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29782 benchmarks/message_tier_bench.py --lanes 16 > gpurun_out/c8_msg_tier_n8.json 2> gpurun_out/c8_msg_tier_n8.err; echo "msg rc=$?"; cat gpurun_out/c8_msg_tier_n8.json | cut -c1-1500; grep -E "Error" gpurun_out/c8_msg_tier_n8.err | head -3
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29783 benchmarks/workloads_bench.py --slots 1000000000 --steps 20 > gpurun_out/c8_workloads_n8_1b.json 2> gpurun_out/c8_workloads_n8_1b.err; echo "workloads rc=$?"; cat gpurun_out/c8_workloads_n8_1b.json | cut -c1-1200
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29784 bench.py --gpus 8 --impl nccl --steps 20 --warmup 3 > gpurun_out/c8_nccl.json 2> gpurun_out/c8_nccl.err; echo "nccl rc=$?"; python -c "import json;d=json.load(open('gpurun_out/c8_nccl.json'));print(d['value']/1e9, d['ms_per_step'])"
