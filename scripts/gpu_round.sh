#!/usr/bin/env bash
# One GPU round: parity tests, smoke, bench at the two batch sizes, ncu launch list + full capture.
# usage (from the repo root, under gpurun): bash scripts/gpu_round.sh [tag]
mkdir -p gpurun_out
TAG="${1:-r1}"
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python bench.py --steps 200 --warmup 10 > gpurun_out/bench.log 2> gpurun_out/bench.err
python bench.py --steps 100 --warmup 10 --envs 65536 --no-cpu-baseline > gpurun_out/bench_65k.log 2>> gpurun_out/bench.err
CMD="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-soak"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:motion_query -s 5 -c 3 -f -o gpurun_out/prof_${TAG}_query $CMD > gpurun_out/ncu_full.log 2>&1
CMD2="python bench.py --steps 6 --warmup 3 --envs 65536 --no-cpu-baseline --no-soak"
$CMD2 > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:motion_query -s 4 -c 2 -f -o gpurun_out/prof_${TAG}_query65k $CMD2 > gpurun_out/ncu_full65k.log 2>&1
python scripts/bench_loss.py > gpurun_out/bench_loss.log 2>&1; tail -1 gpurun_out/bench_loss.log
python scripts/bench_sweep.py > gpurun_out/bench_sweep.log 2>&1; tail -1 gpurun_out/bench_sweep.log
python scripts/bench_motion_opt.py > gpurun_out/bench_motion_opt.log 2>&1; tail -1 gpurun_out/bench_motion_opt.log
python scripts/bench_tracker_step.py > gpurun_out/bench_tracker_step.log 2>&1; tail -1 gpurun_out/bench_tracker_step.log
python scripts/bench_loader.py > gpurun_out/bench_loader.log 2>&1; tail -1 gpurun_out/bench_loader.log
CMD3="python scripts/bench_tracker_step.py --steps 3 --no-cpu"
$CMD3 > gpurun_out/plain4.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}_step.csv $CMD3 > gpurun_out/ncu_list_step.log 2>&1
$CMD3 > gpurun_out/plain5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"sim_step|tar_obs|hf_obs" -s 12 -c 6 -f -o gpurun_out/prof_${TAG}_step $CMD3 > gpurun_out/ncu_full_step.log 2>&1
CMD5="python scripts/bench_loader.py --clips 1024 --host-clips 16"
$CMD5 > gpurun_out/plain6.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:build_tables -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_loader $CMD5 > gpurun_out/ncu_full_loader.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/smoke.log
python - <<'PY'
import json
for f in ("bench.log", "bench_65k.log"):
    try:
        j = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        r = j["roofline"]
        ts = j.get("tracker_step") or {}
        print(f, "value %.3e  ms/step %.4f  frac %.3f  achieved %.0f GB/s  e2e %.3e  cpu %s | step: %.3e bf/s %.4f ms frac %.3f" % (
            j["value"], j["ms_per_step"], r["frac"], r["achieved"], j["e2e"]["value"],
            (j.get("cpu_baseline") or {}).get("value"), ts.get("value", 0), ts.get("ms_per_step", 0), ts.get("roofline_frac", 0)))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -2 gpurun_out/bench.err; tail -2 gpurun_out/ncu_full.log
