#!/bin/bash
# Round-2 ncu evidence (one GPU): full captures of the kernels the design notes quote, and the launch list of the bench.
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on -f"
timeout 600 $N -k regex:sim_scan -s 3 -c 1 -o gpurun_out/r02_ncu_dense_gemm python tools/time_gemm.py 2 > gpurun_out/ncu1.log 2>&1; echo "gemm $?"
timeout 600 $N -k regex:sim_scan -s 2 -c 2 -o gpurun_out/r02_ncu_wide_scan python tools/profile_wide.py > gpurun_out/ncu2.log 2>&1; echo "wide $?"
timeout 600 $N -k regex:"hs_|ranks_transpose64" -c 4 -o gpurun_out/r02_ncu_ranks_hs python tools/time_ranks.py 1 c3 > gpurun_out/ncu3.log 2>&1; echo "ranks $?"
timeout 600 $N -k regex:"clahe_lut|clahe_interp" -s 8 -c 2 -o gpurun_out/r02_ncu_clahe_v10 python tools/time_clahe.py > gpurun_out/ncu4.log 2>&1; echo "clahe $?"
timeout 600 $N -k regex:"topk_finalize|shard_exchange" -s 6 -c 1 -o gpurun_out/r02_ncu_finalize python tools/profile_step.py 125126 > gpurun_out/ncu5.log 2>&1; echo "finalize $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 3 --warmup 3 --cpu-rows 20000 > gpurun_out/ncu_bench.log 2>&1; echo "launch list $?"
tail -2 gpurun_out/ncu1.log gpurun_out/ncu2.log | cut -c1-200
