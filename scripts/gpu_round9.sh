#!/usr/bin/env bash
# usage: gpu_round9.sh N tag  -- tests (N=1 only), bench with the driver's flags on N GPUs
N="${1:-1}"; TAG="${2:-r2k}"
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu_${TAG}.log; tail -3 gpurun_out/pytest_gpu_${TAG}.log
  python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1_${TAG}.log 2> gpurun_out/bench_n1_${TAG}.err; tail -3 gpurun_out/bench_n1_${TAG}.err
else
  python -m pytest tests/test_multi_gpu.py -m gpu -q -x 2>&1 | tail -5
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_${TAG}.log 2> gpurun_out/bench_n${N}_${TAG}.err; tail -5 gpurun_out/bench_n${N}_${TAG}.err
fi
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/bench_n${N}_${TAG}.log").read().strip().splitlines()[-1])
    r = j["roofline"]
    print("value %.3e ms/step %.4f frac %.3f e2e %.3e e2e_sel %.3e" % (j["value"], j["ms_per_step"], r["frac"], j["e2e"]["value"], j["e2e_body_pos_obs_only"]["value"]))
    print("cfg3", {k: v for k, v in j["cfg3"].items() if k not in ("what", "stats", "bound")})
    print("cfg4", json.dumps(j["cfg4"])[:3000]); print("cfg5", {k: v for k, v in j["cfg5"].items() if k not in ("what", "stats")}); print("selfcheck", j["selfcheck"])
except Exception as e:
    print("unreadable", e)
PY
