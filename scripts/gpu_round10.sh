#!/usr/bin/env bash
# tracker step: fused target observation -- parity tests, then the step bench in its configurations
TAG="${1:-r2m}"
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu_${TAG}.log; tail -3 gpurun_out/pytest_gpu_${TAG}.log
python scripts/bench_tracker_step.py --separate-tar-obs --no-cpu > gpurun_out/step_sep_${TAG}.json 2> gpurun_out/step_sep_${TAG}.err; tail -2 gpurun_out/step_sep_${TAG}.err
for v in 0 3 5; do
  python scripts/bench_tracker_step.py --query-variant $v --no-cpu > gpurun_out/step_fused_v${v}_${TAG}.json 2> gpurun_out/step_fused_v${v}_${TAG}.err; tail -2 gpurun_out/step_fused_v${v}_${TAG}.err
done
python scripts/bench_tracker_step.py --separate-tar-obs --query-variant 3 --no-cpu > gpurun_out/step_sep_v3_${TAG}.json 2>/dev/null
python - <<PY
import json
for f in ("step_sep", "step_sep_v3", "step_fused_v0", "step_fused_v3", "step_fused_v5"):
    try:
        j = json.loads(open("gpurun_out/%s_${TAG}.json" % f).read().strip().splitlines()[-1])
        print(f, j["parc_launches_per_step"], "eager %.2f us graph %.2f us frac %.3f" % (j["eager"]["ms_per_step"] * 1e3, j["cuda_graph"]["ms_per_step"] * 1e3, j["cuda_graph"]["roofline_frac"]))
    except Exception as e:
        print(f, "unreadable", e)
PY
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
