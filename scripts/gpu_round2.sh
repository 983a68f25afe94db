#!/usr/bin/env bash
# Round-2 GPU pass: parity tests, smoke, bench (driver's flags), variant sweep.  usage: bash scripts/gpu_round2.sh [tag]
mkdir -p gpurun_out
TAG="${1:-r2}"
python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.log 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python scripts/bench_variants.py --steps 50 > gpurun_out/variants_${TAG}.json 2> gpurun_out/variants_${TAG}.txt; tail -3 gpurun_out/variants_${TAG}.txt
python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/bench_r2.log").read().strip().splitlines()[-1])
    r = j["roofline"]
    print("value %.3e ms/step %.4f frac %.3f serial %.3f flushed %.3f e2e %.3e" % (j["value"], j["ms_per_step"], r["frac"], r["frac_serial"], r["frac_isolated_flushed"], j["e2e"]["value"]))
    print("cfg4", j["cfg4"]); print("cfg5", {k: v for k, v in j["cfg5"].items() if k != "what"}); print("step", j["tracker_step"]["ms_per_step"], j["tracker_step"]["roofline_frac"])
except Exception as e:
    print("unreadable", e)
PY
