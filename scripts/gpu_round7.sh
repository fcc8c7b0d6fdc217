#!/bin/bash
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l); echo "gpus=$NG"
timeout 600 python -m pytest tests/test_gpu_rings.py tests/test_gpu_sketch.py tests/test_gpu_topk.py -x -q --timeout 120 > gpurun_out/pytest_gpu7.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu7.log
show () { python -c "import json,sys;d=json.load(open('$1'));print('$2',d['n_gpus'],'gpus',round(d['value']/1e9,3),'G/s', round(d['ms_per_step'],4),'ms e2e',round(d['e2e']['value']/1e9,3), d['clocks']['reasons'])"; }
if [ $NG -ge 2 ]; then
for v in 0 3 1 5; do
  FPS_MF_REG_VARIANT=$v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/b7_n2_v$v.json 2> gpurun_out/b7_n2_v$v.err; show gpurun_out/b7_n2_v$v.json n2_v$v; grep -i -E "error|Traceback" gpurun_out/b7_n2_v$v.err | head -3
done
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi7.log 2>&1; echo "multi rc=$?"; tail -3 gpurun_out/pytest_multi7.log
fi
