#!/usr/bin/env bash
mkdir -p gpurun_out
TAG="${1:-r2b}"
python -m pytest tests -m gpu -q -x -k "variants or fused_obs or out_of_range or empty_observation" 2>&1 | tail -5
python scripts/bench_variants.py --steps 50 --envs 4096,8192,16384 > gpurun_out/variants_${TAG}.json 2> gpurun_out/variants_${TAG}.txt; tail -2 gpurun_out/variants_${TAG}.txt
python scripts/bench_variants.py --steps 20 --envs 4096 > gpurun_out/variants_${TAG}_k20.json 2> gpurun_out/variants_${TAG}_k20.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.log 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
CMD="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-soak --no-cfg4 --no-cfg5"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:motion_query -s 30 -c 3 -f -o gpurun_out/prof_${TAG}_query $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
