#!/usr/bin/env python
"""Config 5 of BASELINE.json (per-GPU shard): dataset sweep over synthetic clips -- FK on every frame, foot/hand
contact labels, nearest-cell heightfield under every body, per-frame cell masks + per-cell min body height.
Reports body-frames/s (and achieved GB/s for the FK-only pass, 556 algorithmic B per character-frame:
34 floats in, 105 floats out) next to the CPU oracle on a stated sub-sample.

    python scripts/bench_sweep.py [--clips 12500] [--frames 265] [--steps 5]
(12500 clips = one GPU's shard of the 100k-clip config on 8 GPUs.)
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, steps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=12500)
    ap.add_argument("--frames", type=int, default=265)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--cpu-clips", type=int, default=2)
    args = ap.parse_args()
    from parc_b200 import ops
    from parc_b200.anim.kin_char_model import KinCharModel
    from parc_b200.util import geom_util, synth
    from parc_b200.zmotion_editing_tools.motion_edit_lib import label_clips

    dev = torch.device("cuda", 0)
    km = KinCharModel(dev)
    km.load_char_file(os.path.join(ROOT, "parc_b200", "assets", "humanoid.xml"))
    rng = np.random.default_rng(5)
    B, F = args.clips, args.frames
    base_hf = [synth.box_terrain(rng, h_range=(-0.4, 0.7)) if i % 2 else synth.stairs_terrain(rng) for i in range(64)]
    nb = 256                                     # distinct synthetic clips, tiled to B (content does not affect cost)
    fr_np = np.concatenate([synth.synth_clips(km, 4, seed=100 + i, num_frames=F, hf=base_hf[i % 64])[0] for i in range(nb // 4)])
    frames = torch.tensor(fr_np).to(dev).repeat((B + nb - 1) // nb, 1, 1)[:B].contiguous()
    hfs = torch.tensor(np.stack([base_hf[(i // 4) % 64] for i in range(nb)])).to(dev).repeat((B + nb - 1) // nb, 1, 1)[:B].contiguous()
    tb = ops.make_terrain_batch(hfs, torch.zeros(B, 2, device=dev), (0.4, 0.4), base_z=hfs.amin(dim=(1, 2)) - 10.0)
    pts = geom_util.get_char_point_samples(km)
    model = km.c_model()
    flat = frames.reshape(-1, frames.shape[-1])

    ms_fk = timed(lambda: ops.frames_fk(model, flat), args.steps)
    ms_label = timed(lambda: label_clips(frames, tb, km, body_points=None, want_body_hf=True), args.steps)
    ms_full = timed(lambda: label_clips(frames, tb, km, body_points=pts, want_masks=True, want_body_hf=True, want_fk=True),
                    args.steps)
    n_frames = B * F
    # CPU leg (oracle port): lives in bench.py, the one place that may execute oracle/
    import bench
    cpu = bench.cpu_leg_sweep(fr_np, base_hf, args.cpu_clips, F)
    out = {
        "workload": f"cfg5 shard: {B} clips x {F} frames ({n_frames} character-frames), one 16x16 terrain per clip",
        "fk_only": {"ms": ms_fk, "body_frames_per_s": n_frames * 15 / (ms_fk * 1e-3),
                    "achieved_GBps": n_frames * 556 / (ms_fk * 1e-3) / 1e9, "algorithmic_bytes_per_char_frame": 556},
        "fk_contacts_bodyhf": {"ms": ms_label, "body_frames_per_s": n_frames * 15 / (ms_label * 1e-3)},
        "fk_contacts_bodyhf_masks": {"ms": ms_full, "body_frames_per_s": n_frames * 15 / (ms_full * 1e-3)},
        "cpu_oracle": cpu,
    }
    out["speedup_label_vs_cpu"] = out["fk_contacts_bodyhf"]["body_frames_per_s"] / out["cpu_oracle"]["body_frames_per_s_label"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
