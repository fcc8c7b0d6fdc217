#!/bin/bash
# Re-measure the scaling table with the round-1 final code (L2 blocking + single-pass exchange):
#   gpurun --gpus 8 --timeout 900 -- 'bash scripts/next_round_8gpu.sh'      (charged 8x)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
NG=$(nvidia-smi -L | wc -l); echo "gpus=$NG"
run () { n=$1; port=$2; shift 2
  timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port "$@"; }
show () { python -c "import json;d=json.loads(open('$1').read().strip().splitlines()[-1]);print('$2',d['n_gpus'],'gpus',round(d['value']/1e9,3),'G/s',round(d['ms_per_step'],4),'ms e2e',round(d['e2e']['value']/1e9,3),d['clocks']['reasons'])" 2>/dev/null || { echo "$2 FAILED"; tail -4 ${1%.json}.err; }; }
for n in $NG 4; do
  [ $n -le $NG ] || continue
  run $n 2961$n bench.py --gpus $n --steps 200 --warmup 10 > gpurun_out/nr_b_n$n.json 2> gpurun_out/nr_b_n$n.err; show gpurun_out/nr_b_n$n.json default
done
run $NG 29630 bench.py --gpus $NG --steps 200 --warmup 10 --sync-every 8 > gpurun_out/nr_b_n${NG}_k8.json 2> gpurun_out/nr_b_n${NG}_k8.err; show gpurun_out/nr_b_n${NG}_k8.json sync8
run $NG 29631 tests/mp_topk_check.py 2>&1 | grep -E "MP_TOPK_CHECK_OK|Error|assert" | head -3
