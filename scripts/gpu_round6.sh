#!/usr/bin/env bash
# Single-GPU pass: all parity tests, bench (driver flags), TMA-staging variant vs the plain one, loader / sweep benches.
TAG="${1:-r2g}"
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu_${TAG}.log; tail -3 gpurun_out/pytest_gpu_${TAG}.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.log 2> gpurun_out/bench_${TAG}.err; tail -3 gpurun_out/bench_${TAG}.err
python scripts/bench_variants.py --steps 50 --envs 4096,8192 --variants 5,6 --repeat 3 > gpurun_out/variants_${TAG}.json 2> gpurun_out/variants_${TAG}.txt; grep "pdl_early \|serial" gpurun_out/variants_${TAG}.txt
python scripts/bench_loader.py > gpurun_out/loader_${TAG}.json 2> gpurun_out/loader_${TAG}.err; tail -2 gpurun_out/loader_${TAG}.err; cut -c1-400 gpurun_out/loader_${TAG}.json
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/bench_${TAG}.log").read().strip().splitlines()[-1])
    r = j["roofline"]
    print("value %.3e ms/step %.4f frac %.3f serial %.3f flushed %.3f e2e %.3e" % (j["value"], j["ms_per_step"], r["frac"], r["frac_serial"], r["frac_isolated_flushed"], j["e2e"]["value"]))
    print("cfg3", j["cfg3"]); print("cfg4", j["cfg4"]); print("cfg5", {k: v for k, v in j["cfg5"].items() if k not in ("what", "stats")}); print("step", j["tracker_step"]["ms_per_step"], j["tracker_step"]["roofline_frac"])
except Exception as e:
    print("unreadable", e)
PY
