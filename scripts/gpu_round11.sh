#!/usr/bin/env bash
# N-GPU pass: bench under torchrun with the driver's flags + the peer-gather microbenchmark.  usage: gpu_round11.sh N tag
N="${1:-8}"; TAG="${2:-r2n}"
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_${TAG}.log 2> gpurun_out/bench_n${N}_${TAG}.err; tail -5 gpurun_out/bench_n${N}_${TAG}.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/bench_peer.py > gpurun_out/peer_n${N}_${TAG}.json 2> gpurun_out/peer_n${N}_${TAG}.err; tail -3 gpurun_out/peer_n${N}_${TAG}.err
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/bench_n${N}_${TAG}.log").read().strip().splitlines()[-1])
    print("value %.3e frac %.3f e2e %.3e e2e_sel %.3e" % (j["value"], j["roofline"]["frac"], j["e2e"]["value"], j["e2e_body_pos_obs_only"]["value"]))
    print("cfg3", {k: v for k, v in j["cfg3"].items() if k not in ("what", "stats", "bound")})
    c4 = j["cfg4"]; pg = c4.get("peer_gather") or {}
    print("cfg4 shard %.4f eff %.3f nccl %.4f piped %.4f | push %.4f only %.4f piped %.4f p2p %.4f direct %.4f equal %s" % (c4["shard_ms_per_step"], c4["efficiency"], c4["gather"]["ms"], c4["pipelined_ms_per_step"], pg.get("push_ms_per_step", -1), pg.get("push_only_ms_per_step", -1), pg.get("push_pipelined_ms_per_step", -1), pg.get("push_only_peer_pointers_ms_per_step", -1), pg.get("direct_ms_per_step", -1), pg.get("equals_nccl_gather_full_size")))
    print("cfg5", {k: v for k, v in j["cfg5"].items() if k not in ("what", "stats")}); print("selfcheck", j["selfcheck"])
    p = json.loads(open("gpurun_out/peer_n${N}_${TAG}.json").read().strip().splitlines()[-1])
    for r in p["rows"]: print(r)
except Exception as e:
    print("unreadable", e)
PY
