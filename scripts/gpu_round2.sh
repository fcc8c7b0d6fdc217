#!/bin/bash
# 2-GPU contact: fabric + fused step across ranks, N=1/2 bench for fused and NCCL baseline, ncu capture
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l); echo "gpus=$NG"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu2.log
run_bench () { # n impl extra
  if [ "$1" = "1" ]; then timeout 900 python bench.py --gpus 1 --impl $2 $3
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --impl $2 $3; fi
}
for n in 1 2; do
  [ $n -le $NG ] || continue
  run_bench $n fps_b200 "--steps 300 --warmup 10" > gpurun_out/bench_fused_n$n.json 2> gpurun_out/bench_fused_n$n.err; echo "fused n=$n rc=$?"; cat gpurun_out/bench_fused_n$n.json
  run_bench $n nccl "--steps 30 --warmup 5" > gpurun_out/bench_nccl_n$n.json 2> gpurun_out/bench_nccl_n$n.err; echo "nccl n=$n rc=$?"; cat gpurun_out/bench_nccl_n$n.json; tail -3 gpurun_out/bench_nccl_n$n.err
done
# launch list + full capture of the fused kernel (1 GPU)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fps_mf_sgd_fused -s 4 -c 2 -o gpurun_out/prof_mf_fused -f python bench.py --steps 4 --warmup 3 > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out
