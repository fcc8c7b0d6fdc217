#!/bin/bash
# new tests (sampler, length pruning, replica triggers, checkpoint) + pruned top-K benchmark + bench sanity
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_topk.py -x -q -k "sampler or user_memory or length_sorted or replica_exchange or checkpoint or negative_sampling or matches_reference" 2>&1 | tail -8
timeout 200 python benchmarks/topk_bench.py --skew 1.0 > gpurun_out/topk_skew.json 2> gpurun_out/topk_skew.err; cat gpurun_out/topk_skew.json; tail -2 gpurun_out/topk_skew.err
timeout 200 python benchmarks/topk_bench.py > gpurun_out/topk_uniform.json 2> gpurun_out/topk_uniform.err; cat gpurun_out/topk_uniform.json; tail -2 gpurun_out/topk_uniform.err
timeout 200 python bench.py --steps 20 --warmup 5 2> gpurun_out/bench20.err | tee gpurun_out/bench20.json; tail -2 gpurun_out/bench20.err
timeout 200 python bench.py --steps 20 --warmup 5 --format arrays 2> gpurun_out/bench20a.err | tee gpurun_out/bench20a.json
