#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
NG=$(nvidia-smi -L | wc -l); echo "gpus=$NG"
timeout 150 python -m pytest tests/test_gpu_topk.py tests/test_gpu_kernels.py -x -q --timeout 60 --timeout-method=thread > gpurun_out/p11_a.log 2>&1; echo "topk+kernels rc=$?"; tail -3 gpurun_out/p11_a.log
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29617 tests/mp_device_check.py > gpurun_out/p11_mp.log 2>&1; echo "mp rc=$?"; grep -E "OK|Error|error|assert" gpurun_out/p11_mp.log | head -8
show () { python -c "import json,sys;d=json.loads(open('$1').read().strip().splitlines()[-1]);print('$2',d['n_gpus'],'gpus',round(d['value']/1e9,3),'G/s', round(d['ms_per_step'],4),'ms e2e',round(d['e2e']['value']/1e9,3), 'cache',d['config'].get('item_cache'), d['config'].get('sync_every'), d['clocks']['reasons'])" 2>/dev/null || { echo "$2 FAILED"; grep -v "OMP\|\*\*\*" ${1%.json}.err | tail -4; }; }
for k in 1 2 4; do
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 10 --sync-every $k > gpurun_out/b11_n2_k$k.json 2> gpurun_out/b11_n2_k$k.err; show gpurun_out/b11_n2_k$k.json sync$k
done
SWEEP_BYTES=134217728 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 benchmarks/bandwidth_sweep.py > gpurun_out/sweep_n2.jsonl 2> gpurun_out/sweep_n2.err; grep "^{" gpurun_out/sweep_n2.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['dim'], 'pull %.0f GB/s push %.0f GB/s nccl %.0f GB/s'%(d['pull_GBs'],d['push_GBs'],d['nccl_GBs']))"
timeout 200 python benchmarks/workloads_bench.py > gpurun_out/workloads_n1.json 2> gpurun_out/workloads_n1.err; cat gpurun_out/workloads_n1.json; grep -v "OMP\|\*\*\*" gpurun_out/workloads_n1.err | tail -3
