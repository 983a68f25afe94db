#!/usr/bin/env bash
# Loss-kernel pass: all parity tests, cfg3 bench, optimiser bench, ncu of the loss kernel.
TAG="${1:-r2h}"
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu_${TAG}.log; tail -3 gpurun_out/pytest_gpu_${TAG}.log
python scripts/bench_loss.py > gpurun_out/loss_${TAG}.json 2> gpurun_out/loss_${TAG}.err; tail -2 gpurun_out/loss_${TAG}.err; cat gpurun_out/loss_${TAG}.json
python scripts/bench_motion_opt.py > gpurun_out/motion_opt_${TAG}.json 2> gpurun_out/motion_opt_${TAG}.err; tail -2 gpurun_out/motion_opt_${TAG}.err; cat gpurun_out/motion_opt_${TAG}.json
python scripts/bench_sweep.py --cpu-clips 1 > gpurun_out/sweep_${TAG}.json 2> gpurun_out/sweep_${TAG}.err; cut -c1-700 gpurun_out/sweep_${TAG}.json
python scripts/bench_loader.py > gpurun_out/loader_${TAG}.json 2> gpurun_out/loader_${TAG}.err; cut -c1-300 gpurun_out/loader_${TAG}.json
ncu --set full --clock-control none --import-source on -k regex:body_loss -c 2 -f -o gpurun_out/prof_${TAG}_loss python scripts/bench_loss.py --batch 256 --steps 1 --cpu-samples 1 --cpu-frames 2 > gpurun_out/ncu_loss.log 2>&1; tail -1 gpurun_out/ncu_loss.log
