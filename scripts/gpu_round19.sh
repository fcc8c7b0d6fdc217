#!/bin/bash
# ncu captures of the secondary kernels + sanitizer (time-boxed)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:fps_topk_mma -s 2 -c 1 -o gpurun_out/prof_topk_v4 -f python benchmarks/topk_bench.py --queries 2048 --items 400000 > gpurun_out/ncu_topk4.log 2>&1; echo "ncu topk rc=$?"
timeout 200 ncu --set full --clock-control none -k regex:"fps_cache_(push_delta|refresh)" -s 2 -c 2 -o gpurun_out/prof_cache_sync -f python bench.py --steps 8 --warmup 4 --item-cache on --sync-every 1 > gpurun_out/ncu_sync.log 2>&1; echo "ncu sync rc=$?"
timeout 120 python benchmarks/pa_bench.py > gpurun_out/pa_bench.json 2> gpurun_out/pa_bench.err; cat gpurun_out/pa_bench.json; tail -2 gpurun_out/pa_bench.err
timeout 200 ncu --set full --clock-control none -k regex:fps_pa_step -s 3 -c 1 -o gpurun_out/prof_pa -f python benchmarks/pa_bench.py > gpurun_out/ncu_pa.log 2>&1; echo "ncu pa rc=$?"
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_kernels.py -x -q -k "pull_push_dot or (matches_reference and 64) or packed64 or wide_rows or item_cache" > gpurun_out/sanitize_memcheck_kernels.log 2>&1; echo "memcheck kernels rc=$?"; tail -3 gpurun_out/sanitize_memcheck_kernels.log
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_topk.py tests/test_gpu_sketch.py -x -q > gpurun_out/sanitize_memcheck_topk.log 2>&1; echo "memcheck topk/sketch rc=$?"; tail -3 gpurun_out/sanitize_memcheck_topk.log
