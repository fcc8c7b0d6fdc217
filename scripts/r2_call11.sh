#!/bin/bash
# round 2, call 11 (1 GPU): what the driver will run that has not been re-verified since the last changes
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/c11_multi.log 2>&1; echo "multi rc=$? $(tail -1 gpurun_out/c11_multi.log)"
grep -E "Error|assert |FAILED" gpurun_out/c11_multi.log | head -12
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_output.py tests/test_gpu_pa.py -x -q -m gpu > gpurun_out/c11_kernels.log 2>&1; echo "kernels rc=$? $(tail -1 gpurun_out/c11_kernels.log)"
grep -E "Error|assert |FAILED" gpurun_out/c11_kernels.log | head -8
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c11_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/c11_smoke.log)"
