#!/bin/bash
# round 2, call 7 (1 GPU): message-tier optimisations, fp64, top-K pass1_fraction, CTR at 1B slots, NCCL-style baseline at N=1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 python -m pytest tests/test_gpu_rings.py tests/test_gpu_kernels.py -x -q --timeout 150 -k "rings or message or fp64" > gpurun_out/c7_rings.log 2>&1; echo "rings rc=$? $(tail -1 gpurun_out/c7_rings.log)"
grep -E "Error|assert" gpurun_out/c7_rings.log | head -8
for L in 16 64; do timeout 100 python benchmarks/message_tier_bench.py --lanes $L > gpurun_out/c7_msg_n1_l$L.json 2> gpurun_out/c7_msg_n1_l$L.err; echo "msg lanes=$L rc=$?"; python -c "import json;d=json.load(open('gpurun_out/c7_msg_n1_l$L.json'));print([(r['case'], round(r['messages_per_s_per_gpu']/1e6,2)) for r in d['runs']])"; done
timeout 150 python benchmarks/topk_bench.py --pass1-fraction 0.125 > gpurun_out/c7_topk_pass1.json 2> gpurun_out/c7_topk_pass1.err; echo "topk rc=$?"; tail -2 gpurun_out/c7_topk_pass1.json | cut -c1-600
timeout 200 python benchmarks/workloads_bench.py --slots 1000000000 --steps 20 > gpurun_out/c7_workloads_n1_1b.json 2> gpurun_out/c7_workloads_n1_1b.err; echo "workloads rc=$?"; cat gpurun_out/c7_workloads_n1_1b.json | cut -c1-900
timeout 150 python bench.py --impl nccl --steps 20 --warmup 3 > gpurun_out/c7_bench_nccl_n1.json 2> gpurun_out/c7_bench_nccl_n1.err; echo "nccl rc=$?"; python -c "import json;d=json.load(open('gpurun_out/c7_bench_nccl_n1.json'));print(d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9)"
timeout 200 python bench.py --steps 200 --warmup 10 --quality-updates-per-user 0 > gpurun_out/c7_bench_n1.json 2> gpurun_out/c7_bench_n1.err; echo "bench rc=$?"; python -c "import json;d=json.load(open('gpurun_out/c7_bench_n1.json'));print(d['value']/1e9, d['e2e']['value']/1e9, d['value_fp64'])"
