#!/bin/bash
# Multi-GPU evidence for one world size N (run as: gpurun --gpus N -- 'bash tools/gpu_multi_round.sh N'):
# sharded == single-GPU parity log, the per-step overhead breakdown, and the bench line (C4 + C5 at N GPUs).
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29611 tools/test_multi_gpu.py > gpurun_out/r02_multi_gpu_${N}gpu.log 2>&1
echo "multi-gpu parity exit=$? : $(grep -c PASS gpurun_out/r02_multi_gpu_${N}gpu.log) PASS, $(grep -c FAIL gpurun_out/r02_multi_gpu_${N}gpu.log) FAIL"
tail -2 gpurun_out/r02_multi_gpu_${N}gpu.log | cut -c1-200
timeout 300 $TR --master-port 29612 tools/time_sharded_overhead.py > gpurun_out/r02_sharded_overhead_${N}gpu.log 2>&1
tail -6 gpurun_out/r02_sharded_overhead_${N}gpu.log | cut -c1-250
timeout 900 $TR --master-port 29613 bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/r02_bench_${N}gpu_v2.json 2> gpurun_out/r02_bench_${N}gpu_v2.err
echo "bench exit=$?"
tail -c 1500 gpurun_out/r02_bench_${N}gpu_v2.json
