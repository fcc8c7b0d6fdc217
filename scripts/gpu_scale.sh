#!/bin/bash
# N-GPU measurement pass (run under `gpurun --gpus N`): headline bench as the driver runs it, message tier, workloads.
#   gpurun --gpus 8 --timeout 900 -- 'bash scripts/gpu_scale.sh 8'
# A fresh multi-GPU box needs >= 150 s per torchrun stage (first `import torch` + NCCL start-up).
N=${1:-8}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29801 bench.py --gpus $N > gpurun_out/scale_bench_n$N.json 2> gpurun_out/scale_bench_n$N.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/scale_bench_n$N.json'));print(d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d.get('value_direct'), d['config'].get('quality'))"
timeout 240 $TR --master-port 29802 benchmarks/message_tier_bench.py --lanes 64 > gpurun_out/scale_msg_tier_n$N.json 2> gpurun_out/scale_msg_tier_n$N.err; echo "msg rc=$?"; cat gpurun_out/scale_msg_tier_n$N.json | cut -c1-1500
timeout 300 $TR --master-port 29803 benchmarks/workloads_bench.py --slots 1000000000 --steps 20 > gpurun_out/scale_workloads_n$N.json 2> gpurun_out/scale_workloads_n$N.err; echo "workloads rc=$?"; cat gpurun_out/scale_workloads_n$N.json | cut -c1-1200
timeout 240 $TR --master-port 29804 bench.py --gpus $N --impl nccl --steps 20 --warmup 3 > gpurun_out/scale_nccl_n$N.json 2> gpurun_out/scale_nccl_n$N.err; echo "nccl rc=$?"
