#!/usr/bin/env bash
# Final single-GPU pass of the round: all parity tests, smoke, bench (default flags and the driver's), reference arm,
# launch list, full ncu captures of the query / loss / label / loader kernels.
TAG="${1:-r2z}"
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu_${TAG}.log; tail -3 gpurun_out/pytest_gpu_${TAG}.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; tail -1 gpurun_out/smoke_${TAG}.log
python bench.py > gpurun_out/bench_${TAG}.log 2> gpurun_out/bench_${TAG}.err; tail -3 gpurun_out/bench_${TAG}.err
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_k20_${TAG}.log 2> gpurun_out/bench_k20_${TAG}.err
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref_${TAG}.log 2>&1
python scripts/bench_loss.py > gpurun_out/loss_${TAG}.json 2>/dev/null
python scripts/bench_motion_opt.py > gpurun_out/motion_opt_${TAG}.json 2>/dev/null
python scripts/bench_sweep.py > gpurun_out/sweep_${TAG}.json 2>/dev/null
python scripts/bench_loader.py > gpurun_out/loader_${TAG}.json 2>/dev/null
CMD="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-soak --no-cfg5 --no-cfg3"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_list.log 2>&1
CMD2="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-soak --no-cfg5 --no-cfg4 --no-cfg3"
ncu --set full --clock-control none --import-source on -k regex:motion_query -s 30 -c 3 -f -o gpurun_out/prof_${TAG}_query $CMD2 > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
ncu --set full --clock-control none --import-source on -k regex:body_loss -c 2 -f -o gpurun_out/prof_${TAG}_loss python scripts/bench_loss.py --batch 256 --steps 1 --cpu-samples 1 --cpu-frames 2 > gpurun_out/ncu_loss.log 2>&1; tail -1 gpurun_out/ncu_loss.log
ncu --set full --clock-control none --import-source on -k regex:clip_label -s 2 -c 1 -f -o gpurun_out/prof_${TAG}_label_masks python scripts/bench_sweep.py --cpu-clips 1 --steps 1 > gpurun_out/ncu_label.log 2>&1; tail -1 gpurun_out/ncu_label.log
ncu --set full --clock-control none --import-source on -k regex:build_tables -c 2 -f -o gpurun_out/prof_${TAG}_loader python scripts/bench_loader.py > gpurun_out/ncu_loader.log 2>&1; tail -1 gpurun_out/ncu_loader.log
python - <<PY
import json
for f in ("bench_${TAG}", "bench_k20_${TAG}"):
    try:
        j = json.loads(open("gpurun_out/%s.log" % f).read().strip().splitlines()[-1])
        r = j["roofline"]
        print(f, "value %.3e ms/step %.4f frac %.3f serial %.3f flushed %.3f e2e %.3e sel %.3e" % (j["value"], j["ms_per_step"], r["frac"], r["frac_serial"], r["frac_isolated_flushed"], j["e2e"]["value"], j["e2e_body_pos_obs_only"]["value"]))
        print("  cfg3 %.2f ms cfg4 %.4f ms (%.3f) cfg5 %.1f ms step %.4f (%.3f)" % (j["cfg3"]["ms_fwd_bwd"], j["cfg4"]["shard_ms_per_step"], j["cfg4"]["roofline_frac"], j["cfg5"]["ms_per_pass"], j["tracker_step"]["ms_per_step"], j["tracker_step"]["roofline_frac"]))
    except Exception as e:
        print(f, "unreadable", e)
print(open("gpurun_out/bench_ref_${TAG}.log").read().strip().splitlines()[-1][:300])
PY
