#!/usr/bin/env bash
# quick pass: all parity tests + the loss / optimiser benches
TAG="${1:-r2i}"
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu_${TAG}.log; tail -3 gpurun_out/pytest_gpu_${TAG}.log
python scripts/bench_loss.py > gpurun_out/loss_${TAG}.json 2> gpurun_out/loss_${TAG}.err; tail -2 gpurun_out/loss_${TAG}.err; cut -c1-400 gpurun_out/loss_${TAG}.json
python scripts/bench_motion_opt.py > gpurun_out/motion_opt_${TAG}.json 2> gpurun_out/motion_opt_${TAG}.err; tail -2 gpurun_out/motion_opt_${TAG}.err; cut -c1-300 gpurun_out/motion_opt_${TAG}.json
