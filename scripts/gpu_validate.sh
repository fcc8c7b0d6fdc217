#!/bin/bash
# One-GPU validation pass (run under gpurun): every GPU test group with its own timeout, smoke(), N=1 bench.
#   gpurun --timeout 1800 -- 'bash scripts/gpu_validate.sh'
# Logs land in gpurun_out/validate_*.log.  Each stage is bounded: a hang costs minutes, not the whole budget.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; timeout 600 python -m pytest "$@" -x -q > gpurun_out/validate_$name.log 2>&1; echo "$name rc=$? $(tail -1 gpurun_out/validate_$name.log)"; }
run kernels tests/test_gpu_kernels.py tests/test_gpu_output.py tests/test_gpu_pa.py -m gpu
run backend tests/test_device_backend.py
run topk tests/test_gpu_topk.py tests/test_gpu_workloads.py -m gpu
run sketch tests/test_gpu_sketch.py
run rings tests/test_gpu_rings.py
run multi tests/test_gpu_multi.py            # ranks share the GPU when fewer GPUs than ranks are visible
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/validate_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/validate_smoke.log)"
timeout 300 python bench.py > gpurun_out/validate_bench_n1.json 2> gpurun_out/validate_bench_n1.err; echo "bench rc=$?"
grep -hE "^(FAILED|ERROR)|Error" gpurun_out/validate_*.log | head -20
