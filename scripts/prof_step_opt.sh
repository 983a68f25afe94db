#!/usr/bin/env bash
# ncu --set full of the tracker-step kernels and of the optimiser's kernels (one launch each, after the programs ran clean)
mkdir -p gpurun_out
python scripts/bench_tracker_step.py --no-cpu --steps 3 > gpurun_out/plain_step.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"sim_step|tar_obs|hf_obs|motion_query" -s 20 -c 5 -f -o gpurun_out/prof_r2z_step python scripts/bench_tracker_step.py --no-cpu --steps 3 > gpurun_out/ncu_step.log 2>&1; tail -1 gpurun_out/ncu_step.log
python scripts/bench_motion_opt.py > gpurun_out/plain_opt.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"motion_opt|body_loss|frames_fk" -s 12 -c 4 -f -o gpurun_out/prof_r2z_opt python scripts/bench_motion_opt.py > gpurun_out/ncu_opt.log 2>&1; tail -1 gpurun_out/ncu_opt.log
