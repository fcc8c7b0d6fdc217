#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_topk.py -x -q > gpurun_out/pytest_topk.log 2>&1; echo "topk pytest rc=$?"; tail -25 gpurun_out/pytest_topk.log
show () { python -c "import json,sys;d=json.load(open('$1'));print('$2',d['kernel'],round(d['value']/1e9,3),'G/s', round(d['ms_per_step'],4),'ms e2e',round(d['e2e']['value']/1e9,3))"; }
for v in 3 5 6 7 8 9 10; do
  FPS_MF_REG_VARIANT=$v timeout 300 python bench.py --steps 200 --warmup 10 --kernel reg > gpurun_out/b5_reg$v.json 2> gpurun_out/b5_reg$v.err; show gpurun_out/b5_reg$v.json reg$v; tail -2 gpurun_out/b5_reg$v.err
done
