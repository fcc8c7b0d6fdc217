#!/bin/bash
# 8-GPU scaling run: fabric check, MF bench at N=8/4/2 (item-cache on, direct mode for comparison),
# NCCL baseline, bandwidth sweep, w2v/CTR.  Everything time-boxed.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
NG=$(nvidia-smi -L | wc -l); echo "gpus=$NG"
run () { # n port args...
  n=$1; port=$2; shift 2
  timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port "$@"
}
show () { python -c "import json,sys;d=json.loads(open('$1').read().strip().splitlines()[-1]);print('$2',d['n_gpus'],'gpus',round(d['value']/1e9,3),'G/s', round(d['ms_per_step'],4),'ms e2e',round(d['e2e']['value']/1e9,3), 'cache',d['config'].get('item_cache'), d['config'].get('sync_every'), d['clocks']['reasons'])" 2>/dev/null || { echo "$2 FAILED"; grep -v "OMP\|\*\*\*" ${1%.json}.err | tail -4; }; }
run $NG 29617 tests/mp_device_check.py > gpurun_out/p13_mp.log 2>&1; echo "mp rc=$?"; grep -E "OK|rank0.*(Error|assert|Mismatch|Greatest|!=)" gpurun_out/p13_mp.log | head -6
run $NG 29512 bench.py --gpus $NG --steps 200 --warmup 10 > gpurun_out/b13_n${NG}.json 2> gpurun_out/b13_n${NG}.err; show gpurun_out/b13_n${NG}.json default
run $NG 29513 bench.py --gpus $NG --steps 200 --warmup 10 --sync-every 2 > gpurun_out/b13_n${NG}_k2.json 2> gpurun_out/b13_n${NG}_k2.err; show gpurun_out/b13_n${NG}_k2.json sync2
run $NG 29515 bench.py --gpus $NG --steps 100 --warmup 10 --item-cache off > gpurun_out/b13_n${NG}_direct.json 2> gpurun_out/b13_n${NG}_direct.err; show gpurun_out/b13_n${NG}_direct.json direct
run $NG 29516 bench.py --gpus $NG --steps 20 --warmup 3 --impl nccl > gpurun_out/b13_n${NG}_nccl.json 2> gpurun_out/b13_n${NG}_nccl.err; show gpurun_out/b13_n${NG}_nccl.json nccl
for n in 4 2; do
  [ $n -lt $NG ] || continue
  run $n 2952$n bench.py --gpus $n --steps 200 --warmup 10 > gpurun_out/b13_n$n.json 2> gpurun_out/b13_n$n.err; show gpurun_out/b13_n$n.json default
done
SWEEP_BYTES=134217728 run $NG 29530 benchmarks/bandwidth_sweep.py > gpurun_out/sweep_n$NG.jsonl 2> gpurun_out/sweep_n$NG.err; grep "^{" gpurun_out/sweep_n$NG.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['dim'], 'pull %.0f GB/s push %.0f GB/s nccl %.0f GB/s'%(d['pull_GBs'],d['push_GBs'],d['nccl_GBs']))"
run $NG 29531 benchmarks/workloads_bench.py --slots 400000000 > gpurun_out/workloads_n$NG.json 2> gpurun_out/workloads_n$NG.err; grep "^{" gpurun_out/workloads_n$NG.json; grep -v "OMP\|\*\*\*" gpurun_out/workloads_n$NG.err | tail -3
