#!/bin/bash
# round 2, call 12 (1 GPU): ring tests after the fair credit share + final N=1 message-tier numbers
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 250 python -m pytest tests/test_gpu_rings.py -x -q > gpurun_out/c12_rings.log 2>&1; echo "rings rc=$? $(tail -1 gpurun_out/c12_rings.log)"
grep -E "Error|assert |FAILED" gpurun_out/c12_rings.log | head -8
for L in 16 64; do timeout 100 python benchmarks/message_tier_bench.py --lanes $L > gpurun_out/c12_msg_n1_l$L.json 2> gpurun_out/c12_msg_n1_l$L.err; echo "msg lanes=$L rc=$?"; python -c "import json;d=json.load(open('gpurun_out/c12_msg_n1_l$L.json'));print([(r['case'], round(r['messages_per_s_per_gpu']/1e6,2)) for r in d['runs']])"; done
