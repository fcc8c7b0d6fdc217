#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "bucket or blocking or fit_stream or graph" 2>&1 | tail -5
for mode in off on; do
  timeout 200 python bench.py --steps 20 --warmup 5 --item-blocking $mode 2> gpurun_out/bench23_$mode.err | tee gpurun_out/bench23_$mode.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$mode', d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['gpu_launches'], d['config']['item_blocking'])"
done
timeout 120 python benchmarks/sketch_bench.py > gpurun_out/sketch_bench.json 2> gpurun_out/sketch_bench.err; cat gpurun_out/sketch_bench.json; tail -2 gpurun_out/sketch_bench.err
