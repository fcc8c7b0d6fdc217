#!/bin/bash
# round 2, call 9 (1 GPU): lane-per-message message tier, credit-counter trims, pass1_fraction default, CTR at 1B slots
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 250 python -m pytest tests/test_gpu_rings.py -x -q --timeout 200 > gpurun_out/c9_rings.log 2>&1; echo "rings rc=$? $(tail -1 gpurun_out/c9_rings.log)"
grep -E "Error|assert|FAILED" gpurun_out/c9_rings.log | head -8
for L in 16 64; do timeout 100 python benchmarks/message_tier_bench.py --lanes $L > gpurun_out/c9_msg_n1_l$L.json 2> gpurun_out/c9_msg_n1_l$L.err; echo "msg lanes=$L rc=$?"; python -c "import json;d=json.load(open('gpurun_out/c9_msg_n1_l$L.json'));print([(r['case'], round(r['messages_per_s_per_gpu']/1e6,2)) for r in d['runs']])"; grep -E "Error" gpurun_out/c9_msg_n1_l$L.err | head -3; done
timeout 300 python -m pytest tests/test_device_backend.py tests/test_gpu_topk.py tests/test_gpu_workloads.py -x -q --timeout 150 -m gpu > gpurun_out/c9_misc.log 2>&1; echo "misc rc=$? $(tail -1 gpurun_out/c9_misc.log)"
grep -E "Error|assert|FAILED" gpurun_out/c9_misc.log | head -8
timeout 200 python benchmarks/workloads_bench.py --slots 1000000000 --steps 20 > gpurun_out/c9_workloads_n1_1b.json 2> gpurun_out/c9_workloads_n1_1b.err; echo "workloads rc=$?"; python -c "import json;d=json.load(open('gpurun_out/c9_workloads_n1_1b.json'));print(d['ctr'])"
