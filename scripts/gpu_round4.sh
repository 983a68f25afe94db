#!/usr/bin/env bash
# Dataset-sweep / loader pass: parity tests of the labelling + loader kernels, their benches, ncu captures.
# usage: bash scripts/gpu_round4.sh [tag]
mkdir -p gpurun_out
TAG="${1:-r2d}"
python -m pytest tests -m gpu -q -k "label or mask or loader or packed or frames_fk or quirk" 2>&1 | tail -15 > gpurun_out/pytest_sweep_${TAG}.log
tail -3 gpurun_out/pytest_sweep_${TAG}.log
python scripts/bench_sweep.py --cpu-clips 1 > gpurun_out/sweep_${TAG}.json 2> gpurun_out/sweep_${TAG}.err; tail -2 gpurun_out/sweep_${TAG}.err; cat gpurun_out/sweep_${TAG}.json
python scripts/bench_loader.py > gpurun_out/loader_${TAG}.json 2> gpurun_out/loader_${TAG}.err; tail -2 gpurun_out/loader_${TAG}.err; cat gpurun_out/loader_${TAG}.json
ncu --set full --clock-control none --import-source on -k regex:clip_label -c 2 -f -o gpurun_out/prof_${TAG}_label python scripts/bench_sweep.py --cpu-clips 1 --steps 1 > gpurun_out/ncu_label.log 2>&1; tail -1 gpurun_out/ncu_label.log
ncu --set full --clock-control none --import-source on -k regex:build_tables -c 2 -f -o gpurun_out/prof_${TAG}_loader python scripts/bench_loader.py > gpurun_out/ncu_loader.log 2>&1; tail -1 gpurun_out/ncu_loader.log
