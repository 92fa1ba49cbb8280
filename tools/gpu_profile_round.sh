#!/bin/bash
# ncu evidence for the bench step: launch list of the bench command itself + full capture of the dominant kernel.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -k regex:"sim_scan|select_kth|topk_finalize|rescore|pack_bf16|elementwise|fill" -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 3 --warmup 3 --no-extras --cpu-rows 20000 > gpurun_out/ncu_bench.log 2>&1
echo "launch list exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:sim_scan -s 4 -c 2 -f -o gpurun_out/prof_sim_scan_v11 python tools/profile_step.py --steps 6 > gpurun_out/ncu_full.log 2>&1
echo "full exit=$?"
tail -3 gpurun_out/ncu_bench.log | cut -c1-300
tail -3 gpurun_out/ncu_full.log | cut -c1-300
