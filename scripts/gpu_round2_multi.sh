#!/usr/bin/env bash
# N-GPU pass: NCCL test, bench under torchrun (driver's flags), D2H microbenchmark.  usage: gpu_round2_multi.sh N
N="${1:-2}"
mkdir -p gpurun_out
nvidia-smi -L | head -8
python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}.log 2> gpurun_out/bench_n${N}.err; tail -3 gpurun_out/bench_n${N}.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/bench_d2h.py > gpurun_out/d2h_n${N}.json 2> gpurun_out/d2h_n${N}.err; tail -2 gpurun_out/d2h_n${N}.err; cat gpurun_out/d2h_n${N}.json
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/bench_n${N}.log").read().strip().splitlines()[-1])
    print("value %.3e frac %.3f e2e %.3e" % (j["value"], j["roofline"]["frac"], j["e2e"]["value"]))
    print("cfg4", json.dumps(j["cfg4"])); print("cfg5", {k: v for k, v in j["cfg5"].items() if k not in ("what","stats")}); print("selfcheck", j["selfcheck"])
except Exception as e:
    print("unreadable", e)
PY
