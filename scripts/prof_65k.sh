#!/usr/bin/env bash
# ncu --set full of the 65 536-env launch (cfg4 at N = 1) and the launch list of one tracker step
mkdir -p gpurun_out
CMD="python bench.py --envs 65536 --steps 10 --warmup 3 --no-cpu-baseline --no-soak --no-cfg5 --no-cfg4 --no-cfg3"
$CMD > gpurun_out/plain65.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:motion_query -s 20 -c 2 -f -o gpurun_out/prof_r2z_query65k $CMD > gpurun_out/ncu65.log 2>&1; tail -1 gpurun_out/ncu65.log
python scripts/bench_tracker_step.py --no-cpu --steps 3 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sim_step|tar_obs|hf_obs|motion_query" --csv --log-file gpurun_out/launches_r2_tracker_step.csv python scripts/bench_tracker_step.py --no-cpu --steps 3 > /dev/null 2>&1
tail -3 gpurun_out/launches_r2_tracker_step.csv | cut -c1-200
