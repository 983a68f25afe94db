#!/usr/bin/env python
"""Tuning sweep of motion_query_kernel: every instantiation (include/parc_b200.h: ParcQueryArgs.variant) x launch form
(serial graph / programmatic dependent launch) x batch size, K back-to-back steps as one CUDA graph over distinct
input batches (bench.py's headline method).  Prints one JSON object; used to choose the regime thresholds in
csrc/motion_query.cu and recorded under profiles/.

    python scripts/bench_variants.py [--steps 50] [--envs 4096,8192,16384,32768,65536]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--envs", default="2048,4096,8192,16384,32768,65536")
    ap.add_argument("--clips", type=int, default=2048)
    ap.add_argument("--variants", default="0,1,2,3,4,5,6")
    ap.add_argument("--repeat", type=int, default=1, help="repeat the whole sweep (box-to-box / run-to-run noise)")
    a = ap.parse_args()
    args = argparse.Namespace(envs=4096, clips=a.clips, steps=a.steps, warmup=3, no_pdl=False)
    ctx = bench.setup(args)
    ctx.peak = 6555.5
    out = {"steps": a.steps, "rows": []}
    NB = 16
    for envs in [int(x) for x in a.envs.split(",")] * a.repeat:
        ids_h, times_h = bench.query_batches(NB, envs, a.clips, 264.0 / 30.0, seed=7)
        ids_d, times_d = ids_h.to(ctx.dev), times_h.to(ctx.dev)
        for variant in [int(v) for v in a.variants.split(",")]:
            for mode in ("serial", "pdl", "pdl_early", "pdl_early_fast_heading"):
                kw = dict(variant=variant, pdl=mode != "serial", pdl_early_inputs=mode.startswith("pdl_early"),
                          fast_heading=mode.endswith("fast_heading"))
                res = {}
                plans = bench.make_plans(ctx, ids_d, times_d, res, **kw)
                ms = bench.timed_graph_steps(ctx, plans, a.steps, 3) / a.steps
                frac = envs * bench.BYTES_PER_CHAR_FRAME / (ms * 1e-3) / 1e9 / ctx.peak
                out["rows"].append({"envs": envs, "variant": variant, "mode": mode, "us": ms * 1e3, "frac": frac})
                print(f"envs {envs:6d} variant {variant} {mode:24s} {ms * 1e3:8.2f} us  frac {frac:.3f}", file=sys.stderr)
                del plans, res
    print(json.dumps(out))


if __name__ == "__main__":
    main()
