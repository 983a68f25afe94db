#!/usr/bin/env python
"""building / opening a motion library.  Host build (the reference's per-clip torch-CPU chain, what
`MotionLib(..., build_on_device=False)` runs), the one-launch GPU loader, and opening a packed `.parcpack` file.

    python scripts/bench_loader.py [--clips 2048] [--frames 265]
"""
import argparse
import json
import os
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=2048)
    ap.add_argument("--frames", type=int, default=265)
    ap.add_argument("--host-clips", type=int, default=256, help="clips timed through the host build (scaled linearly)")
    args = ap.parse_args()
    import __graft_entry__ as entry
    entry.ensure_built()
    from parc_b200 import ops
    from parc_b200.anim.kin_char_model import KinCharModel
    from parc_b200.anim.motion_lib import LoopMode, MotionLib
    from parc_b200.util import synth

    dev = torch.device("cuda", 0)
    km = KinCharModel(dev)
    km.load_char_file(os.path.join(ROOT, "parc_b200", "assets", "humanoid.xml"))
    frames, contacts = synth.synth_clips(km, args.clips, seed=5, num_frames=args.frames)
    M, F = frames.shape[0], frames.shape[1]
    fr_d, ct_d = torch.from_numpy(frames).to(dev), torch.from_numpy(contacts).to(dev)
    torch.cuda.synchronize()

    def wall(fn, reps=3):
        best = 1e9
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = fn()
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        return best, out

    kw = dict(init_type="motion_frames", loop_mode=LoopMode.CLAMP, fps=30, contact_info=True)
    t_gpu, lib = wall(lambda: MotionLib(fr_d, km, dev, contacts=ct_d, **kw))
    # the launch alone
    nf = torch.full((M,), F, dtype=torch.long, device=dev)
    ones = torch.ones(M, device=dev)
    flat, cflat = fr_d.reshape(M * F, -1), ct_d.reshape(M * F, -1)
    model = km.c_model()
    ops.build_tables(model, flat, cflat, nf, 30 * ones, 30 * ones)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.build_tables(model, flat, cflat, nf, 30 * ones, 30 * ones)
    e1.record()
    torch.cuda.synchronize()
    ms_kernel = e0.elapsed_time(e1) / 10
    hc = min(args.host_clips, M)
    t_host, _ = wall(lambda: MotionLib(torch.from_numpy(frames[:hc]), km, dev, contacts=torch.from_numpy(contacts[:hc]), **kw),
                     reps=1)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "lib.parcpack")
        t_save, _ = wall(lambda: lib.save_packed(path), reps=1)
        size = os.path.getsize(path)
        t_open, back = wall(lambda: MotionLib(path, km, dev, init_type="packed_file", contact_info=True))
        same = bool(torch.equal(back._packed.rows, lib._packed.rows))
    lay = lib._packed.layout
    bytes_algo = M * F * ((6 + km.get_dof_size() + km.get_num_joints()) * 4 + lay.row_floats * 4)
    out = {
        "workload": f"{M} clips x {F} frames ({M * F} frames, packed table {M * F * lay.row_floats * 4 / 1e6:.0f} MB)",
        "gpu_loader": {"s_total": t_gpu, "ms_launch": ms_kernel, "frames_per_s_launch": M * F / (ms_kernel * 1e-3),
                       "achieved_GBps": bytes_algo / (ms_kernel * 1e-3) / 1e9,
                       "algorithmic_bytes_per_frame": bytes_algo // (M * F)},
        "host_build": {"s_measured": t_host, "clips_measured": hc, "s_scaled_to_library": t_host * M / hc,
                       "cores": os.cpu_count()},
        "packed_file": {"bytes": size, "s_save": t_save, "s_open": t_open, "rows_identical": same},
        "speedup_gpu_loader_vs_host": (t_host * M / hc) / t_gpu,
        "speedup_open_vs_host": (t_host * M / hc) / t_open,
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
