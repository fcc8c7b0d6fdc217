#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q > gpurun_out/pytest_gpu4.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu4.log
show () { python -c "import json,sys;d=json.load(open('$1'));print('$2',d['kernel'],round(d['value']/1e9,3),'G/s', round(d['ms_per_step'],4),'ms e2e',round(d['e2e']['value']/1e9,3))"; }
timeout 300 python bench.py --steps 200 --warmup 10 --kernel tma > gpurun_out/b4_tma.json 2> gpurun_out/b4_tma.err; show gpurun_out/b4_tma.json tma; tail -2 gpurun_out/b4_tma.err
for v in 0 1 2 3 4; do
  FPS_MF_REG_VARIANT=$v timeout 300 python bench.py --steps 200 --warmup 10 --kernel reg > gpurun_out/b4_reg$v.json 2> gpurun_out/b4_reg$v.err; show gpurun_out/b4_reg$v.json reg$v; tail -2 gpurun_out/b4_reg$v.err
done
