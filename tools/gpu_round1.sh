#!/bin/bash
# One gpurun call: bench line, clocks, ncu launch list, ncu --set full on the dominant kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv &
SMI=$!
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit=$?"
kill $SMI
tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_search.csv python tools/profile_step.py --steps 4 > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sim_scan -s 4 -c 2 -f -o gpurun_out/prof_sim_scan python tools/profile_step.py --steps 4 > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit=$?"
ls -la gpurun_out/
