#!/usr/bin/env bash
# Round-2 full pass: all GPU parity tests, smoke, bench with the driver's flags, launch list, one full ncu capture.
# usage: bash scripts/gpu_round3.sh [tag]
mkdir -p gpurun_out
TAG="${1:-r2c}"
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu_${TAG}.log
tail -3 gpurun_out/pytest_gpu_${TAG}.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; tail -1 gpurun_out/smoke_${TAG}.log
python bench.py > gpurun_out/bench_${TAG}.log 2> gpurun_out/bench_${TAG}.err; tail -3 gpurun_out/bench_${TAG}.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_${TAG}.log 2>&1
CMD="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-soak --no-cfg5"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_list.log 2>&1
CMD2="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-soak --no-cfg5 --no-cfg4"
$CMD2 > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:motion_query -s 30 -c 3 -f -o gpurun_out/prof_${TAG}_query $CMD2 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/bench_${TAG}.log").read().strip().splitlines()[-1])
    r = j["roofline"]
    print("value %.3e ms/step %.4f frac %.3f serial %.3f flushed %.3f e2e %.3e" % (j["value"], j["ms_per_step"], r["frac"], r["frac_serial"], r["frac_isolated_flushed"], j["e2e"]["value"]))
    print("cfg4", j["cfg4"]); print("cfg5", {k: v for k, v in j["cfg5"].items() if k not in ("what", "stats")}); print("step", j["tracker_step"]["ms_per_step"], j["tracker_step"]["roofline_frac"])
except Exception as e:
    print("unreadable", e)
PY
