#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2> gpurun_out/bench25_n2.err | tee gpurun_out/bench25_n2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=2', round(d['value']/1e9,3), round(d['ms_per_step'],4), round(d['e2e']['value']/1e9,3), d['config']['item_blocking'], d['config']['item_cache'])"
tail -2 gpurun_out/bench25_n2.err
