#!/usr/bin/env bash
# N-GPU pass for the peer-memory gather: NCCL + peer test, bench under torchrun (cfg4 legs), plus the sweep / loader benches.
# usage: bash scripts/gpu_round5.sh N [tag]
N="${1:-2}"
TAG="${2:-r2e}"
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -12
python -m pytest tests/test_multi_gpu.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_multi_${TAG}.log; tail -8 gpurun_out/pytest_multi_${TAG}.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 ${BENCH_EXTRA:-} > gpurun_out/bench_n${N}_${TAG}.log 2> gpurun_out/bench_n${N}_${TAG}.err; tail -5 gpurun_out/bench_n${N}_${TAG}.err
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/bench_n${N}_${TAG}.log").read().strip().splitlines()[-1])
    print("value %.3e frac %.3f e2e %.3e" % (j["value"], j["roofline"]["frac"], j["e2e"]["value"]))
    print("cfg4", json.dumps(j["cfg4"])); print("selfcheck", j["selfcheck"])
except Exception as e:
    print("unreadable", e)
PY
[ -n "${SKIP_SINGLE:-}" ] && exit 0
python scripts/bench_sweep.py --cpu-clips 1 > gpurun_out/sweep_${TAG}.json 2> gpurun_out/sweep_${TAG}.err; tail -2 gpurun_out/sweep_${TAG}.err; cat gpurun_out/sweep_${TAG}.json
python scripts/bench_loader.py > gpurun_out/loader_${TAG}.json 2> gpurun_out/loader_${TAG}.err; tail -2 gpurun_out/loader_${TAG}.err; cat gpurun_out/loader_${TAG}.json
