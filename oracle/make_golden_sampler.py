"""Golden vectors of the MDM sampler's per-sample terrain gather from the REAL reference, and the oracle pin.

TEST INFRASTRUCTURE; runs only in the authoring container (needs /root/reference):
    python -m oracle.make_golden_sampler
MDMHeightfieldContactMotionSampler.get_hfs_from_data (diffusion/mdm_heightfield_contact_motion_sampler.py:449-474) and
its helper (:414-447) are called unbound on a stand-in object carrying exactly the attributes they read (the sampler's
constructor needs an MDM config and a dataset); augmentation off.  Three clips with their own terrains of different
sizes (50 x 50, 102 x 102, 16 x 16), body-cover masks from the reference's compute_hf_mask_inds
(util/terrain_util.py:1951-1997), both relative-z styles.  Writes tests/golden/sampler_golden.npz.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

ref_shim.activate()
import anim.kin_char_model as ref_kcm  # noqa: E402
import diffusion.mdm_heightfield_contact_motion_sampler as ref_sampler  # noqa: E402
import util.geom_util as ref_geom  # noqa: E402
import util.terrain_util as ref_terrain  # noqa: E402
import util.torch_util as ref_tu  # noqa: E402

from oracle import parc_oracle as O  # noqa: E402

REPORT = []


def pin(name, ref, mine):
    ok = ref.shape == mine.shape and torch.equal(ref, mine)
    REPORT.append(f"{'OK ' if ok else 'FAIL'} {name} shape={tuple(ref.shape)}")
    if not ok:
        print("\n".join(REPORT))
        raise SystemExit(f"oracle != reference for {name}")


def main():
    rng = np.random.default_rng(7)
    torch.manual_seed(7)
    km = ref_kcm.KinCharModel("cpu")
    km.load_char_file(os.path.join(ref_shim.REFERENCE_ROOT, "data/assets/humanoid.xml"))
    body_points = ref_geom.get_char_point_samples(km)
    civ = np.load(os.path.join(GOLD, "clip_civilization.npz"))
    tea = np.load(os.path.join(GOLD, "clip_teaser_terrain.npz"))
    NF = 24                                               # frames per clip that carry masks

    def terrain(hf, min_point, dxdy, name):
        t = ref_terrain.SubTerrain(name, x_dim=hf.shape[0], y_dim=hf.shape[1], dx=float(dxdy[0]), dy=float(dxdy[1]),
                                   min_x=float(min_point[0]), min_y=float(min_point[1]), device="cpu")
        t.hf = torch.tensor(hf).clone()
        band = rng.uniform(0.1, 0.8, size=hf.shape + (2,)).astype(np.float32)
        t.hf_maxmin = torch.stack([t.hf + torch.tensor(band[..., 0]), t.hf - torch.tensor(band[..., 1])], dim=-1)
        return t

    fr_civ = torch.tensor(civ["frames"][100:100 + NF])
    t0 = terrain(civ["hf"], civ["min_point"], civ["dxdy"], "civ")
    # clip 1: the same motion moved onto the 102 x 102 teaser terrain; clip 2: on a 16 x 16 box terrain
    t1 = terrain(tea["hf"], tea["min_point"], tea["dxdy"], "teaser")
    fr1 = fr_civ.clone()
    fr1[:, 0:2] += torch.tensor(tea["min_point"]) + 0.4 * 50 - fr_civ[0, 0:2]
    fr1[:, 2] += 0.3
    small = np.zeros((16, 16), dtype=np.float32)
    for _ in range(6):
        x0, y0 = rng.integers(0, 12, size=2)
        small[x0:x0 + rng.integers(2, 5), y0:y0 + rng.integers(2, 5)] = np.float32(rng.uniform(-0.5, 0.8))
    t2 = terrain(small, np.array([-3.2, -3.2], np.float32), np.array([0.4, 0.4], np.float32), "small")
    fr2 = fr_civ.clone()
    fr2[:, 0:2] -= fr_civ[NF // 2, 0:2]
    terrains, clips = [t0, t1, t2], [fr_civ, fr1, fr2]
    mask_inds = []
    for t, fr in zip(terrains, clips):
        inds, _ = ref_terrain.compute_hf_mask_inds(fr, t, km, body_points)
        mask_inds.append(inds)

    # samples: (clip, frame window) with the root taken from the window's reference frame, slightly perturbed
    B, T = 40, 5
    ids = torch.tensor(rng.integers(0, 3, size=B))
    starts = rng.integers(0, NF - T + 1, size=B)
    mti = torch.tensor(np.stack([np.arange(s, s + T) for s in starts]))
    root_pos = torch.stack([clips[int(c)][int(s) + 1, 0:3] for c, s in zip(ids, starts)]) + torch.tensor(
        rng.uniform(-0.3, 0.3, size=(B, 3)).astype(np.float32))
    root_rot = ref_tu.exp_map_to_quat(torch.stack([clips[int(c)][int(s) + 1, 3:6] for c, s in zip(ids, starts)]) + torch.tensor(
        rng.uniform(-0.5, 0.5, size=(B, 3)).astype(np.float32)))
    canon_z = root_pos[:, 2].clone()
    num_neg, num_pos, dx, max_h = 15, 15, 0.2, 3.0          # diffusion/mdm.yaml:137-143
    zero = torch.zeros(2)
    grid = ref_geom.get_xy_grid_points(zero, dx, dx, num_neg, num_pos, num_neg, num_pos)

    out = dict(ids=ids.numpy(), mti=mti.numpy(), root_pos=root_pos.numpy(), root_rot=root_rot.numpy(), canon_z=canon_z.numpy(),
               grid=grid.numpy(), num_neg=np.int32(num_neg), dx=np.float32(dx), max_h=np.float32(max_h), num_frames=np.int32(NF))
    for c, (t, fr, inds) in enumerate(zip(terrains, clips, mask_inds)):
        out[f"hf{c}"], out[f"maxmin{c}"] = t.hf.numpy(), t.hf_maxmin.numpy()
        out[f"min_point{c}"], out[f"dxdy{c}"] = t.min_point.numpy(), t.dxdy.numpy()
        out[f"mask_count{c}"] = np.array([i.shape[0] for i in inds])
        out[f"mask_inds{c}"] = torch.cat(inds, dim=0).numpy()

    o_terr = [O.Terrain(hf=t.hf, min_point=t.min_point, dxdy=t.dxdy) for t in terrains]
    for style in (ref_sampler.RelativeZStyle.RELATIVE_TO_ROOT_FLOOR, ref_sampler.RelativeZStyle.RELATIVE_TO_ROOT):
        cls = ref_sampler.MDMHeightfieldContactMotionSampler
        fake = types.SimpleNamespace(
            _generic_heightmap=grid, _grid_dim_x=2 * num_neg + 1, _grid_dim_y=2 * num_neg + 1, _num_x_neg=num_neg,
            _num_y_neg=num_neg, _relative_z_style=style, _use_hf_augmentation=False, _device="cpu", _max_h=max_h,
            _min_h=-max_h, _mlib=types.SimpleNamespace(_terrains=terrains, _hf_mask_inds=mask_inds))
        captured = {}

        def helper(motion_ids, xy_points, motion_time_indices, _f=fake, _c=captured):
            hfs, mms = cls.get_hfs_from_data_helper(_f, motion_ids, xy_points, motion_time_indices)
            _c["mm"] = torch.stack(mms, dim=0)
            return hfs, mms

        fake.get_hfs_from_data_helper = helper
        hfs, center_h = cls.get_hfs_from_data(fake, ids, root_pos, root_rot, canon_z, mti)
        tag = style.name.lower()
        ref_rel = center_h if style == ref_sampler.RelativeZStyle.RELATIVE_TO_ROOT_FLOOR else canon_z
        mm = captured["mm"] - ref_rel.unsqueeze(-1).unsqueeze(-1).unsqueeze(-1)      # :466 / :472
        o_hfs, o_ch, o_mm = O.clip_hfs_from_data(o_terr, [t.hf_maxmin for t in terrains], mask_inds, ids, root_pos, root_rot,
                                                 canon_z, mti, grid, num_neg, num_neg, max_h,
                                                 style == ref_sampler.RelativeZStyle.RELATIVE_TO_ROOT)
        pin(f"get_hfs_from_data[{tag}] hfs", hfs, o_hfs)
        pin(f"get_hfs_from_data[{tag}] center_h", center_h, o_ch)
        pin(f"get_hfs_from_data[{tag}] hf_maxmins", mm, o_mm)
        out[f"hfs_{tag}"], out[f"center_h_{tag}"], out[f"mm_{tag}"] = hfs.numpy(), center_h.numpy(), mm.numpy()
        assert (captured["mm"][..., 0] < max_h * 2.0).float().mean() > 0.02, "masks must cover some sampled cells"
    # pre-rounding grid coordinates of every sampled point, so tests can identify cell-border samples
    heading = ref_tu.calc_heading(root_rot).unsqueeze(-1).unsqueeze(-1).expand(-1, 31, 31)
    xy = ref_tu.rotate_2d_vec(grid.unsqueeze(0).expand(B, -1, -1, -1), heading) + root_pos[:, 0:2].unsqueeze(1).unsqueeze(1)
    out["grid_coord"] = torch.stack([(xy[i] - terrains[int(ids[i])].min_point) / terrains[int(ids[i])].dxdy for i in range(B)]).numpy()
    np.savez_compressed(os.path.join(GOLD, "sampler_golden.npz"), **out)
    with open(os.path.join(GOLD, "PIN_REPORT_sampler.txt"), "w") as f:
        f.write("\n".join(REPORT) + "\n")
    print("\n".join(REPORT))


if __name__ == "__main__":
    main()
