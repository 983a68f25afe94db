"""Golden vectors for SURVEY 8(a) row a16 from the REAL reference, and the pin of the oracle restatement.

TEST INFRASTRUCTURE; runs only in the authoring container (needs /root/reference):
    python -m oracle.make_golden_a16
The MDM's heightfield-collision loss: util/terrain_util.py:1895-1949 (motion_frames_hf_sdf_loss) and the sdf the
MDM builds itself (diffusion/mdm.py:978-1028 compute_point_hf_sdf) with `0.5 * sum(clamp(sdf, max=0)^2)`
(:735, :1493) back-propagated through util/terrain_util.py:1835-1893, on the MDM's 31 x 31 @ 0.2 m local grid
(diffusion/mdm.yaml:137-143) with geom_util.get_minimal_char_point_samples (util/geom_util.py:873-932).
Writes tests/golden/a16_golden.npz and appends to tests/golden/PIN_REPORT_a16.txt.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

ref_shim.activate()
import anim.kin_char_model as ref_kcm  # noqa: E402
import util.geom_util as ref_geom  # noqa: E402
import util.terrain_util as ref_terrain  # noqa: E402
import util.torch_util as ref_tu  # noqa: E402

from oracle import parc_oracle as O  # noqa: E402

REPORT = []


def pin(name, ref, mine, exact=True, rtol=0.0):
    ref, mine = torch.as_tensor(ref).detach(), torch.as_tensor(mine).detach()
    if exact:
        ok = ref.shape == mine.shape and torch.equal(ref, mine)
    else:
        ok = ref.shape == mine.shape and (ref - mine).abs().max().item() <= rtol * ref.abs().max().item()
    REPORT.append(f"{'OK ' if ok else 'FAIL'} {name} shape={tuple(ref.shape)}" + ("" if exact else f" (rtol {rtol:g} norm-wise)"))
    if not ok:
        print("\n".join(REPORT))
        raise SystemExit(f"oracle != reference for {name}: {(ref.double() - mine.double()).abs().max().item()}")


def main():
    torch.manual_seed(0)
    rng = np.random.default_rng(16)
    km = ref_kcm.KinCharModel("cpu")
    km.load_char_file(os.path.join(ref_shim.REFERENCE_ROOT, "data/assets/humanoid.xml"))
    om = O.CharModel.from_npz(os.path.join(GOLD, "humanoid_model.npz"))
    pts_min = ref_geom.get_minimal_char_point_samples(km)
    B, S, X = 4, 15, 31
    dxdy = torch.tensor([0.2, 0.2])
    min_center = (-dxdy * torch.tensor([15.0, 15.0])).unsqueeze(0).expand(B, 2).contiguous()       # mdm.py:1023-1025
    # local heightfields: flat ground with raised / sunken boxes so that feet, hands and some limbs penetrate
    hf = np.zeros((B, X, X), dtype=np.float32)
    for b in range(B):
        for _ in range(10):
            lx, ly = rng.integers(2, 9, size=2)
            x0, y0 = rng.integers(0, X - lx), rng.integers(0, X - ly)
            hf[b, x0:x0 + lx, y0:y0 + ly] = np.float32(rng.uniform(-0.6, 0.9))
    hf = torch.tensor(hf)
    # motions in the heightfield's (character-local) frame: root near the origin, walking forward, ~0.8 m high
    lo = km._lower_dof_limits.numpy() if hasattr(km, "_lower_dof_limits") else -np.ones(28)
    hi = km._upper_dof_limits.numpy() if hasattr(km, "_upper_dof_limits") else np.ones(28)
    t = np.arange(S)[None, :, None] / 30.0
    frames = np.zeros((B, S, 34), dtype=np.float64)
    frames[..., 0:1] = rng.uniform(-0.5, 0.5, size=(B, 1, 1)) + 1.2 * t
    frames[..., 1:2] = rng.uniform(-0.5, 0.5, size=(B, 1, 1)) + 0.3 * t
    frames[..., 2:3] = rng.uniform(0.55, 0.95, size=(B, 1, 1)) + 0.05 * np.sin(6.0 * t)
    frames[..., 3:6] = rng.uniform(-0.4, 0.4, size=(B, 1, 3)) + 0.2 * np.sin(3.0 * t + rng.uniform(0, 6, size=(B, 1, 3)))
    mid, amp = 0.5 * (lo + hi), 0.3 * (hi - lo)
    frames[..., 6:] = mid + amp * np.sin(2 * np.pi * rng.uniform(0.3, 1.5, size=(B, 1, 28)) * t + rng.uniform(0, 6, size=(B, 1, 28)))
    frames[..., 3:][np.abs(frames[..., 3:]) < 1e-3] = 1e-3          # the reference's exp-map gradient is NaN at 0 (F8d)
    frames = torch.tensor(frames, dtype=torch.float32)

    out = {"frames": frames.numpy(), "hf": hf.numpy(), "min_center": min_center.numpy(), "dxdy": dxdy.numpy()}
    out["minimal_point_count"] = np.array([p.shape[0] for p in pts_min])
    out["minimal_points"] = torch.cat(pts_min, dim=0).numpy()
    opts = [torch.tensor(out["minimal_points"][s:s + n]) for s, n in zip(np.cumsum([0] + list(out["minimal_point_count"][:-1])),
                                                                        out["minimal_point_count"])]

    # ---- motion_frames_hf_sdf_loss, interior and exterior distance, value + gradient of the sum ----
    for interior in (True, False):
        tag = "int" if interior else "ext"
        mf = frames.clone().requires_grad_(True)
        loss, wpts, sdf = ref_terrain.motion_frames_hf_sdf_loss(mf, pts_min, hf, min_center, dxdy, km, ret_vis_info=True,
                                                                interior_distance=interior)
        loss.sum().backward()
        mo = frames.clone().requires_grad_(True)
        oloss, opts_w, osdf = O.motion_frames_hf_sdf_loss(om, mo, opts, hf, min_center, dxdy, interior_distance=interior)
        oloss.sum().backward()
        pin(f"motion_frames_hf_sdf_loss[{tag}] loss", loss, oloss)
        pin(f"motion_frames_hf_sdf_loss[{tag}] points", wpts, opts_w)
        pin(f"motion_frames_hf_sdf_loss[{tag}] sdf", sdf, osdf)
        pin(f"motion_frames_hf_sdf_loss[{tag}] d/d motion_frames", mf.grad, mo.grad, exact=False, rtol=2e-6)
        out[f"loss_{tag}"] = loss.detach().numpy()
        out[f"sdf_{tag}"] = sdf.detach().numpy()
        out[f"grad_frames_{tag}"] = mf.grad.numpy()
        if interior:
            out["world_points"] = wpts.detach().numpy()
            assert (sdf < 0).sum().item() > 20, "fixture must contain penetrating points"

    # ---- the MDM's own use: sdf of given world points, loss 0.5*sum(clamp(sdf,max=0)^2), gradient wrt the points ----
    for base_z, tag in ((-10.0, "guidance"), (-5.6, "train")):       # mdm.py:1489 default base_z ; :733 min_h - 5
        p = torch.tensor(out["world_points"]).clone().requires_grad_(True)
        sdf = ref_terrain.points_hf_sdf(p, hf, min_center, dxdy, base_z=base_z)
        l = 0.5 * torch.sum(torch.square(torch.clamp(sdf, max=0.0)))
        l.backward()
        po = torch.tensor(out["world_points"]).clone().requires_grad_(True)
        osdf = O.points_hf_sdf(po, hf, min_center, dxdy, base_z=base_z, inverted=True)
        O.hf_collision_loss(osdf).sum().backward()
        pin(f"points_hf_sdf[{tag}] sdf", sdf, osdf)
        pin(f"points_hf_sdf[{tag}] d loss / d points", p.grad, po.grad)
        out[f"psdf_{tag}"] = sdf.detach().numpy()
        out[f"pgrad_{tag}"] = p.grad.numpy()
        out[f"base_z_{tag}"] = np.float32(base_z)
    # exterior (solid) distance with an upstream gradient of ones, to pin the non-inverted VJP as well
    p = torch.tensor(out["world_points"]).clone().requires_grad_(True)
    sdf = ref_terrain.points_hf_sdf(p, hf, min_center, dxdy, base_z=-10.0, inverted=False)
    w = torch.tensor(rng.uniform(-1, 1, size=tuple(sdf.shape)).astype(np.float32))
    (sdf * w).sum().backward()
    out["psdf_solid"] = sdf.detach().numpy()
    out["pgrad_solid"] = p.grad.numpy()
    out["upstream_solid"] = w.numpy()

    np.savez_compressed(os.path.join(GOLD, "a16_golden.npz"), **out)
    with open(os.path.join(GOLD, "PIN_REPORT_a16.txt"), "w") as f:
        f.write("\n".join(REPORT) + "\n")
    print("\n".join(REPORT))
    print("wrote a16_golden.npz", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
