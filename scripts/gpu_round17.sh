#!/bin/bash
# cold-start timing of the GPU test-suite exactly as the driver runs it
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( time timeout 900 python -m pytest tests/ -x -q -m gpu --durations=8 ) > gpurun_out/p17_full.log 2>&1; echo "rc=$?"; grep -E "passed|failed|real|s call|s setup" gpurun_out/p17_full.log | head -14
