#!/bin/bash
# ncu captures of the kernels added late in the round + the L2-blocked MF step
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 240 $NCU -k regex:"fps_mf_sgd_fused|fps_bucket" -s 12 -c 3 -o gpurun_out/prof_mf_blocked python bench.py --steps 4 --warmup 3 > gpurun_out/ncu_mf_blocked.log 2>&1; echo "ncu mf blocked rc=$?"
timeout 240 $NCU -k regex:fps_cache_exchange -s 2 -c 1 -o gpurun_out/prof_cache_exchange python bench.py --steps 4 --warmup 3 --item-cache on --sync-every 1 > gpurun_out/ncu_exchange.log 2>&1; echo "ncu exchange rc=$?"
timeout 240 $NCU -k regex:fps_pa_step_warp -s 3 -c 1 -o gpurun_out/prof_pa_warp python benchmarks/pa_bench.py > gpurun_out/ncu_pa_warp.log 2>&1; echo "ncu pa rc=$?"
timeout 240 $NCU -k regex:"fps_row_(kth|topk)|fps_topk_mma" -s 17 -c 4 -o gpurun_out/prof_topk_v6 python benchmarks/topk_bench.py --items 400000 > gpurun_out/ncu_topk6.log 2>&1; echo "ncu topk rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -5
