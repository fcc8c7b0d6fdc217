#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu6.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu6.log
show () { python -c "import json,sys;d=json.load(open('$1'));print('$2',d['kernel'],round(d['value']/1e9,3),'G/s', round(d['ms_per_step'],4),'ms e2e',round(d['e2e']['value']/1e9,3), d['clocks'])"; }
timeout 300 python bench.py --steps 300 --warmup 10 > gpurun_out/b6_default.json 2> gpurun_out/b6_default.err; show gpurun_out/b6_default.json default; tail -2 gpurun_out/b6_default.err
timeout 300 python bench.py --steps 300 --warmup 10 --format arrays > gpurun_out/b6_arrays.json 2> gpurun_out/b6_arrays.err; show gpurun_out/b6_arrays.json arrays; tail -2 gpurun_out/b6_arrays.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fps_mf_sgd_fused -s 4 -c 1 -o gpurun_out/prof_mf_final -f python bench.py --steps 4 --warmup 3 > gpurun_out/ncu_final.log 2>&1; echo "ncu rc=$?"
