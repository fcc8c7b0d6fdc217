#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu3.log
for k in tma reg; do
  timeout 300 python bench.py --steps 300 --warmup 10 --kernel $k > gpurun_out/bench_k_$k.json 2> gpurun_out/bench_k_$k.err; echo "kernel=$k rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/bench_k_$k.json'));print(d['kernel'],d['value'],d['ms_per_step'],d['e2e']['value'],d['clocks'])"
  tail -3 gpurun_out/bench_k_$k.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fps_mf_sgd_tma -s 4 -c 1 -o gpurun_out/prof_mf_tma -f python bench.py --steps 4 --warmup 3 --kernel tma > gpurun_out/ncu_tma.log 2>&1; echo "ncu rc=$?"
