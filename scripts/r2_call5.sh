#!/bin/bash
# round 2, call 5 (2 GPUs): N=2 bench with the sliced exchange + own shard in place; CTA / stage sweep; real-NVLink tests
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { # name, env..., -- args
  name=$1; shift
  timeout 300 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29751 bench.py --gpus 2 --steps 200 --warmup 10 $BARGS > gpurun_out/c5_$name.json 2> gpurun_out/c5_$name.err
  echo "$name rc=$? $(python -c "import json;d=json.load(open('gpurun_out/c5_$name.json'));print(round(d['value']/1e9,2),'G', round(d['ms_per_step'],4),'ms e2e',round(d['e2e']['value']/1e9,2), 'direct', d.get('value_direct') and round(d['value_direct']['value']/1e9,2), 'exch', d['config'].get('exchange'))" 2>&1 | tail -1)"
}
BARGS="--quality-updates-per-user 0" run default FPS_EXCHANGE_TIMING=1
BARGS="--no-direct --quality-updates-per-user 0" run ctas48 FPS_EXCHANGE_CTAS=48 FPS_EXCHANGE_TIMING=1
BARGS="--no-direct --quality-updates-per-user 0" run ctas64 FPS_EXCHANGE_CTAS=64 FPS_EXCHANGE_TIMING=1
BARGS="--no-direct --quality-updates-per-user 0" run c16s8 FPS_EXCHANGE_CTAS=16 FPS_EXCHANGE_STAGES=8 FPS_EXCHANGE_TIMING=1
BARGS="--no-direct --quality-updates-per-user 0 --sync-every 2" run se2 X=1
BARGS="--no-direct --quality-updates-per-user 0 --sync-every 1" run se1 X=1
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q --timeout 280 > gpurun_out/c5_multi.log 2>&1; echo "multi rc=$?"; tail -3 gpurun_out/c5_multi.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29752 benchmarks/message_tier_bench.py > gpurun_out/c5_msg_tier_n2.json 2> gpurun_out/c5_msg_tier_n2.err; echo "msg rc=$?"; cat gpurun_out/c5_msg_tier_n2.json; tail -3 gpurun_out/c5_msg_tier_n2.err
