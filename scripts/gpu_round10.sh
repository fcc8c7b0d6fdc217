#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
NG=$(nvidia-smi -L | wc -l); echo "gpus=$NG"
timeout 120 python -m pytest tests/test_gpu_topk.py -x -q --timeout 60 --timeout-method=thread > gpurun_out/p10_topk.log 2>&1; echo "topk rc=$?"; tail -3 gpurun_out/p10_topk.log
timeout 120 python benchmarks/topk_bench.py --queries 2048 > gpurun_out/topk_bench2.json 2> gpurun_out/topk_bench2.err; cat gpurun_out/topk_bench2.json; tail -2 gpurun_out/topk_bench2.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29617 tests/mp_device_check.py > gpurun_out/p10_mp.log 2>&1; echo "mp rc=$?"; grep -E "OK|Error|error|assert" gpurun_out/p10_mp.log | head -8
show () { python -c "import json,sys;d=json.load(open('$1'));print('$2',d['n_gpus'],'gpus',round(d['value']/1e9,3),'G/s', round(d['ms_per_step'],4),'ms e2e',round(d['e2e']['value']/1e9,3), 'cache',d['config'].get('item_cache'), d['clocks']['reasons'])" 2>/dev/null || tail -3 ${1%.json}.err; }
for c in on off; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 10 --item-cache $c > gpurun_out/b10_n2_$c.json 2> gpurun_out/b10_n2_$c.err; show gpurun_out/b10_n2_$c.json cache_$c
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 200 --warmup 10 --sync-every 2 > gpurun_out/b10_n2_s2.json 2> gpurun_out/b10_n2_s2.err; show gpurun_out/b10_n2_s2.json sync2
SWEEP_BYTES=134217728 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 benchmarks/bandwidth_sweep.py > gpurun_out/sweep_n2.jsonl 2> gpurun_out/sweep_n2.err; cat gpurun_out/sweep_n2.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['dim'], 'pull %.0f GB/s push %.0f GB/s nccl %.0f GB/s'%(d['pull_GBs'],d['push_GBs'],d['nccl_GBs']))"; tail -2 gpurun_out/sweep_n2.err
