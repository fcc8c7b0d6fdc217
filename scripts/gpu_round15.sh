#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29617 tests/mp_device_check.py > gpurun_out/p15_mp.log 2>&1; echo "mp rc=$?"; grep -E "OK|rank0.*(Error|assert|Mismatch|Greatest|!=)" gpurun_out/p15_mp.log | head -8
show () { python -c "import json,sys;d=json.loads(open('$1').read().strip().splitlines()[-1]);print('$2',d['n_gpus'],'gpus',round(d['value']/1e9,3),'G/s', round(d['ms_per_step'],4),'ms e2e',round(d['e2e']['value']/1e9,3), 'cache',d['config'].get('item_cache'), d['config'].get('sync_every'), 'mse', d['e2e']['last_step_mse'], d['clocks']['reasons'])" 2>/dev/null || { echo "$2 FAILED"; grep -v "OMP\|\*\*\*" ${1%.json}.err | tail -4; }; }
for k in 1 2 4; do
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$k bench.py --gpus 2 --steps 200 --warmup 10 --sync-every $k > gpurun_out/b15_n2_k$k.json 2> gpurun_out/b15_n2_k$k.err; show gpurun_out/b15_n2_k$k.json sync$k
done
