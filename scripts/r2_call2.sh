#!/bin/bash
# round 2, call 2 (2 GPUs): real-NVLink multi-rank tests, N=2 bench with the new exchange + sweeps,
# single-process peer exchange bench + ncu capture with nvlink counters
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/c2_multi.log 2>&1; echo "multi rc=$?"; tail -4 gpurun_out/c2_multi.log
run() { # name, env..., -- args
  name=$1; shift
  timeout 400 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 2 --steps 200 --warmup 10 $BARGS > gpurun_out/c2_$name.json 2> gpurun_out/c2_$name.err
  echo "$name rc=$? $(python -c "import json;d=json.load(open('gpurun_out/c2_$name.json'));print(round(d['value']/1e9,2),'G', round(d['ms_per_step'],4),'ms e2e',round(d['e2e']['value']/1e9,2), 'direct', d.get('value_direct') and round(d['value_direct']['value']/1e9,2), 'q', d['config'].get('quality'))" 2>&1 | tail -1)"
}
BARGS="" run default X=1
BARGS="--no-direct --quality-updates-per-user 0" run ctas16 FPS_EXCHANGE_CTAS=16
BARGS="--no-direct --quality-updates-per-user 0" run ctas64 FPS_EXCHANGE_CTAS=64
BARGS="--no-direct --quality-updates-per-user 0" run st8 FPS_EXCHANGE_STAGES=8
BARGS="--no-direct --quality-updates-per-user 0 --sync-every 2" run se2 X=1
BARGS="--no-direct --quality-updates-per-user 0 --sync-every 8" run se8 X=1
timeout 200 python benchmarks/exchange_peer_bench.py > gpurun_out/c2_peer_exchange.json 2> gpurun_out/c2_peer_exchange.err; echo "peer rc=$?"; cat gpurun_out/c2_peer_exchange.json
timeout 300 ncu --set full --metrics nvlrx__bytes.sum,nvltx__bytes.sum,lts__t_bytes.sum --clock-control none --import-source on -k regex:fps_replica_exchange -c 2 -o gpurun_out/prof_replica_exchange python benchmarks/exchange_peer_bench.py --once --ctas 32,296 > gpurun_out/c2_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/c2_ncu.log
